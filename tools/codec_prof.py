"""Run one codec decode (+encode) of a few 10 s waveforms on the tiny model — target for ncu captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam_audio_b200.model import build_synthetic_model
from sam_audio_b200.synthetic import synthetic_clip

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = build_synthetic_model("sam-audio-tiny", seed=0, weights_device="cuda")
eng = m._ensure_engine()
lat = torch.randn(n, 250, 256, device="cuda")
wav = torch.empty(n, 2, 480000, device="cuda")
clips = torch.stack([synthetic_clip(i) for i in range(n)]).cuda()
for it in range(2):
    eng.profile(True)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    feats = m._get_audio_features(clips)
    e1.record()
    eng.decode(lat, n, 250, wav)
    e2.record()
    torch.cuda.synchronize()
    rep = eng.profile_report()
    eng.profile(False)
print(f"encode {n} clips: {e0.elapsed_time(e1):.2f} ms; decode {2*n} waveforms: {e1.elapsed_time(e2):.2f} ms")
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:28s} n={int(v['launches']):3d} ms={v['ms']:8.3f} TF/s={v['flops']/max(v['ms'],1e-9)/1e9:8.1f}")
