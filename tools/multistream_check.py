import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
trainer = bench.build_trainer(dev, False)
model, st = trainer.model, trainer.store
img, aud = bench.synth_inputs(64, 1000, False)
img, aud = img.to(dev), aud.to(dev)
def run():
    st.zero_grad()
    torch.manual_seed(7)
    li, la, _, _ = model(img, aud)
    trainer.backward(li + la); trainer.accums = 0
    torch.cuda.synchronize()
    return st.flat_g.clone(), float(li), float(la)
os.environ["DAVF_STREAMS"] = "0"
ref, li0, la0 = run()
ref2, _, _ = run()
print("single-stream repeatability rel", float((ref - ref2).norm() / ref.norm()), "finite", bool(torch.isfinite(ref).all()))
os.environ["DAVF_STREAMS"] = "1"
bad_seen = 0
for it in range(40):
    g, li, la = run()
    rel = float((g - ref).norm() / ref.norm()) if bool(torch.isfinite(g).all()) else float("nan")
    if not (rel < 1e-3):
        bad_seen += 1
        print(f"iter {it}: rel {rel} loss {li} {la} (ref {li0} {la0})")
        for k, name in enumerate(st.names):
            b, e = st.span(k)
            a, r = g[b:e], ref[b:e]
            fin = bool(torch.isfinite(a).all())
            d = float((a - r).norm() / (r.norm() + 1e-20)) if fin else float("nan")
            if not (d < 2e-2):
                print(f"    {name}: rel {d} finite {fin}")
        if bad_seen >= 2:
            break
print("done, bad iterations:", bad_seen)
