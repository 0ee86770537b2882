"""Single-scene (sync) host path: ms per scene by sub-band count, against the device-resident scene and against the same
sub-bands computed from device-resident rasters (what the cuts alone cost)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vitcnn_b200
from vitcnn_b200 import scene as S

dev = "cuda:0"
H, W, C1, C2, P, K = 349, 1905, 144, 1, 11, 16
torch.manual_seed(0)
net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K).to(dev).eval()
img1 = torch.rand(H, W, C1).pin_memory(); img2 = torch.rand(H, W, C2).pin_memory()
d1, d2 = img1.to(dev), img2.to(dev)
lg = torch.empty(H, W, K).pin_memory(); am = torch.empty(H, W, dtype=torch.uint8).pin_memory()

def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("device-resident scene: %.2f ms" % timed(lambda: net.predict_scene(d1, d2)))
for pl in (1, 2, 3, 4, 6):
    S._PLAN.clear()
    os.environ["VITCNN_PIPELINE"] = str(pl)
    ms = timed(lambda: vitcnn_b200.predict_scene_host(net, img1, img2, logits_out=lg, argmax_out=am, sync=True))
    plan = list(S._PLAN.values())
    print("sync, pipeline <= %d: %.2f ms  plan %s" % (pl, ms, plan[0][1] if plan else None))
os.environ.pop("VITCNN_PIPELINE")
# the planned cuts on device-resident rasters (no copies): what the cuts themselves cost
S._PLAN.clear()
vitcnn_b200.predict_scene_host(net, img1, img2, logits_out=lg, argmax_out=am, sync=True)
spans = list(S._PLAN.values())[0][1]
from vitcnn_b200.utils import band_geometry
geo = band_geometry(H, W, P, 1, 0, 1)
xs = geo["xs"]
def cut():
    for a, b in spans:
        xa, xb = int(xs[a]), int(xs[b - 1]) + P
        net.predict_scene(d1[xa:xb], d2[xa:xb], xs=xs[a:b] - xa)
print("planned cuts %s from device-resident rasters: %.2f ms" % (spans, timed(cut)))
