import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
trainer = bench.build_trainer(dev, False)
img, aud = bench.synth_inputs(64, 1000, False)
img, aud = img.to(dev), aud.to(dev)
torch.manual_seed(2000)
NSTEP = int(os.environ.get('NSTEP', '6'))
for i in range(NSTEP):
    li, la, _, _ = trainer.model(img, aud)
    norm, _ = trainer.step(li + la)
    torch.cuda.synchronize()
    st = trainer.store
    if i % max(1, NSTEP // 6) == 0 or not bool(torch.isfinite(norm)):
        print(i, float(li), float(la), float(norm), "p finite", bool(torch.isfinite(st.flat_p).all()), flush=True)
    if not bool(torch.isfinite(norm)):
        break
if len(sys.argv) > 1:
    from deepavfusion_b200.util.graphed import GraphedTrainStep
    g = GraphedTrainStep(trainer, img, aud, warmup=2, capture_error_mode=sys.argv[1])
    for i in range(6):
        li, la, norm = g(img, aud)
        torch.cuda.synchronize()
        print("graph", i, float(li), float(la), float(norm), bool(torch.isfinite(trainer.store.flat_p).all()), flush=True)
