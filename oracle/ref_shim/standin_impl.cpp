// ORACLE / TEST INFRASTRUCTURE ONLY.  The out-of-line part of the Eigen stand-in (matrix product,
// inverse, stream output), compiled once with optimisation; see Eigen/Dense in this directory.
#define EPIVO_ORACLE_EIGEN_STANDIN_IMPL
#include <Eigen/Dense>
