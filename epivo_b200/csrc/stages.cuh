// Declarations of the per-stage launchers shared by api.cu and seq.cu.
#pragma once
#include "common.cuh"
