"""-m gpu, needs >= 2 GPUs (skipped otherwise): data-parallel gradients over NCCL equal the single-GPU
full-batch gradients; ranks stay bit-identical after an optimizer step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import model_utils as U
    from oracle import avmae_oracle as O
    from deepavfusion_b200.util import distributed as D
    from deepavfusion_b200.util.misc import Trainer
    D.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2 * world)
    ni, na = U.make_noise(cfg, 2 * world)
    sl = slice(2 * rank, 2 * rank + 2)
    model = U.build_model(cfg, dev); model.load_state_dict(O.build_state(cfg, seed=rank))     # broadcast must equalise
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95))
    trainer = Trainer(model, optimizer=opt, accum_iter=1, distributed=True, bucket_mb=0.25)
    with U.inject_rand([ni[sl], na[sl]]):
        li, la, _, _ = trainer.model(image[sl].to(dev), audio[sl].to(dev))
    trainer.backward(li + la)
    torch.cuda.synchronize()
    grads = (trainer.store.flat_g * float(trainer.optimizer.scal[2])).cpu()
    trainer.optimizer.step()
    torch.cuda.synchronize()
    torch.save({"grads": grads, "p": trainer.store.flat_p.cpu(), "names": trainer.store.names}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_data_parallel_matches_full_batch(tmp_path):
    import model_utils as U
    from oracle import avmae_oracle as O
    from deepavfusion_b200.util.misc import Trainer
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["grads"], r1["grads"]) and torch.equal(r0["p"], r1["p"])
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 4)
    ni, na = U.make_noise(cfg, 4)
    model = U.build_model(cfg, "cuda"); model.load_state_dict(O.build_state(cfg, seed=0))
    trainer = Trainer(model, optimizer=torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95)))
    with U.inject_rand([ni, na]):
        li, la, _, _ = model(image.cuda(), audio.cuda())
    trainer.backward(li + la)
    torch.cuda.synchronize()
    full = trainer.store.flat_g.cpu()
    rel = ((full - r0["grads"]).norm() / full.norm()).item()
    assert rel < 2e-3, rel


def _graph_worker(rank, world, port, out_dir, overlap):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import model_utils as U
    from oracle import avmae_oracle as O
    from deepavfusion_b200.util import distributed as D
    from deepavfusion_b200.util.misc import Trainer
    from deepavfusion_b200.util.graphed import GraphedTrainStep
    D.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2 * world)
    sl = slice(2 * rank, 2 * rank + 2)
    model = U.build_model(cfg, dev); model.load_state_dict(O.build_state(cfg, seed=0))
    trainer = Trainer(model, optimizer=torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95)), distributed=True, bucket_mb=0.25)
    img, aud = image[sl].to(dev), audio[sl].to(dev)
    torch.manual_seed(7 + rank)
    g = GraphedTrainStep(trainer, img, aud, warmup=1, overlap_comm=overlap)
    assert g.overlap_comm == overlap
    p0 = trainer.store.flat_p.clone()
    torch.manual_seed(100 + rank)                       # same mask noise in both modes
    for _ in range(2):
        li, la, norm = g(img, aud)
    torch.cuda.synchronize()
    torch.save({"dp": (trainer.store.flat_p - p0).cpu(), "loss": float(li) + float(la), "norm": float(norm)}, os.path.join(out_dir, f"g{int(overlap)}_r{rank}.pt"))
    os._exit(0)                                         # no NCCL teardown with a live graph holding NCCL nodes (see bench.py)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_graph_captured_allreduce_matches_post_replay_allreduce(tmp_path):
    """The step graph with the bucketed NCCL all-reduce (and bucketed AdamW) captured inside it produces the same
    parameter update as capturing forward + backward only and all-reducing after the replay; ranks stay identical."""
    world = 2
    for overlap in (True, False):
        ctx = mp.spawn(_graph_worker, args=(world, _free_port(), str(tmp_path), overlap), nprocs=world, join=False)
        for p in ctx.processes:
            p.join(300)
            assert p.exitcode == 0, p.exitcode
    a0, a1 = torch.load(tmp_path / "g1_r0.pt"), torch.load(tmp_path / "g1_r1.pt")
    b0 = torch.load(tmp_path / "g0_r0.pt")
    assert torch.equal(a0["dp"], a1["dp"])
    assert float((a0["dp"] - b0["dp"]).norm()) <= 2e-2 * float(b0["dp"].norm())
    assert abs(a0["norm"] - b0["norm"]) <= 1e-2 * b0["norm"]
