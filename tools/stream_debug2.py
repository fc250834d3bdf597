import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
trainer = bench.build_trainer(dev, False)
img, aud = bench.synth_inputs(64, 1000, False)
img, aud = img.to(dev), aud.to(dev)
def step(bwd, opt):
    li, la, _, _ = trainer.model(img, aud)
    if bwd:
        trainer.backward(li + la)
    if opt:
        trainer.optimizer.step(zero_grad=True, sync_hp=False); trainer.accums = 0
    return li.detach()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step(True, True)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
for name, bwd, opt, mode in [("fwd", False, False, "global"), ("fwd+bwd global", True, False, "global"), ("fwd+bwd relaxed", True, False, "relaxed"), ("fwd+bwd thread_local", True, False, "thread_local"), ("full relaxed", True, True, "relaxed")]:
    try:
        g = torch.cuda.CUDAGraph()
        with torch.no_grad() if not bwd else torch.enable_grad():
            with torch.cuda.graph(g, capture_error_mode=mode):
                out = step(bwd, opt)
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name, float(out), flush=True)
        trainer.store.zero_grad()
    except Exception as e:
        print("FAIL", name, str(e).splitlines()[0][:200], flush=True)
        torch.cuda.synchronize()
