"""e2e probe: per-step wall time, copy time under load, for the async host path."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fishbirdeyevisualslam_b200 import synth, _lib
from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline, PinnedBuffer
B = 128
pipe = FrontBirdPipeline(B)
fr = synth.cheap_batch(B, 720, 1280, 100); bi = synth.cheap_batch(B, 384, 384, 200)
hF, hB = PinnedBuffer(fr.shape), PinnedBuffer(bi.shape)
hF.array[...] = fr; hB.array[...] = bi
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 2
res = [PinnedBuffer((B,), _lib.PAIR_RESULT_DTYPE) for _ in range(3)]
fm = [PinnedBuffer((B, pipe.front_cap), np.int32) for _ in range(3)]
bm = [PinnedBuffer((B, pipe.bird_cap), np.int32) for _ in range(3)]
def loop(n, with_matches=True):
    q = []; cms = []
    t0 = time.perf_counter()
    for i in range(n):
        k = i % 3
        q.append(pipe.submit_host(hF.ptr, hB.ptr, res[k].array, fm[k].array if with_matches else None, bm[k].array if with_matches else None))
        if len(q) >= DEPTH:
            t = q.pop(0)
            pipe.wait(t)
            ms = C.c_float(); pipe._L.fbe_pipeline_copy_ms(pipe._h, t, C.byref(ms)); cms.append(ms.value)
    for t in q:
        pipe.wait(t)
    dt = time.perf_counter() - t0
    return dt / n * 1e3, float(np.median(cms))
loop(5)
for wm in (True, False):
    ms, c = loop(30, wm)
    print(f"depth {DEPTH} with_matches={wm}: {ms:.3f} ms/step wall, median input copy {c:.3f} ms ({(fr.nbytes+bi.nbytes)/c/1e6:.1f} GB/s)")
dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()
for _ in range(3): pipe.step_dev(dF.data_ptr(), dB.data_ptr())
pipe.sync(); t0 = time.perf_counter()
for _ in range(30): pipe.step_dev(dF.data_ptr(), dB.data_ptr())
pipe.sync(); print(f"device-resident loop: {(time.perf_counter()-t0)/30*1e3:.3f} ms/step wall")
