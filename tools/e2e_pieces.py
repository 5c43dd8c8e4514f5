"""GPU tuning helper: e2e time of epivo_seq_process on the benchmark sequence for different upload piece schedules
(EPIVO_UPLOAD_DIV: first piece = one matcher wave / div, doubling; EPIVO_UPLOAD_PIECES: number of pieces)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from epivo_b200 import api, synth
seq = synth.make_sequence(4541, 2000, seed=synth.seed_for(3, 0))
h_kps = torch.from_numpy(seq.kps).pin_memory(); h_desc = torch.from_numpy(seq.descs).pin_memory()
kps, desc = h_kps.numpy(), h_desc.numpy()
ctx = api.Context(0)
pipe = api.SequencePipeline(seq.n_frames, 2000, ctx=ctx)
prm = api.default_params(seq.K.astype(np.float32))
h_res = torch.zeros(seq.n_pairs * api.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
res = h_res.numpy().view(api.RESULT_DTYPE)
for div, pieces in [(1, 4), (2, 5), (4, 6), (4, 5), (8, 7), (1, 3), (2, 4), (1, 5)]:
    os.environ["EPIVO_UPLOAD_DIV"] = str(div); os.environ["EPIVO_UPLOAD_PIECES"] = str(pieces)
    for _ in range(2): pipe.process(prm, kps, desc, res)
    t0 = time.perf_counter()
    for _ in range(5): pipe.process(prm, kps, desc, res)
    dt = (time.perf_counter() - t0) / 5
    print(f"div {div} pieces {pieces}: {dt * 1e3:.3f} ms  {seq.n_pairs / dt:.0f} pairs/s")
