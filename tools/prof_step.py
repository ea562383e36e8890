"""Profiling driver: `python tools/prof_step.py [batch] [steps]` runs warm-up steps of the front+bird pipeline, then
`steps` steps bracketed by cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).  Prints device ms/step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pipe = FrontBirdPipeline(B)
scene = sys.argv[3] if len(sys.argv) > 3 else "survey"          # "survey" = the SURVEY recipe of bench.py, "road" = road-scene statistics
mk = synth.road_batch if scene == "road" else synth.cheap_batch
dF = torch.from_numpy(mk(B, 720, 1280, 100)).cuda()
dB = torch.from_numpy(mk(B, 384, 384, 200)).cuda()
for _ in range(3):
    pipe.step_dev(dF.data_ptr(), dB.data_ptr())
pipe.sync()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
stream = torch.cuda.ExternalStream(pipe.stream_ptr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    pipe.step_dev(dF.data_ptr(), dB.data_ptr())
pipe.join()
e1.record(stream)
pipe.sync()
tot = e0.elapsed_time(e1)
torch.cuda.cudart().cudaProfilerStop()
res, _, _ = pipe.fetch(with_matches=False)
print(f"batch {B}: {tot / steps:.3f} ms/step, {B * steps / tot * 1e3:.0f} pairs/s; mean kps front {res['n_front'].mean():.1f} bird {res['n_bird'].mean():.1f}; "
      f"matches front {res['front_matches'].mean():.1f} bird {res['bird_matches'].mean():.1f}")
