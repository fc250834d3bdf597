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
