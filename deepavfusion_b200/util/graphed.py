"""Whole-step CUDA graph: forward + backward (+ bucket all-reduce) + fused AdamW captured once and
replayed, removing the ~2,000 per-step launches' CPU cost and every host sync from the loop
(eager PyTorch on this path is launch-bound).  Hyper-parameters (lr / weight decay,
beta^t, grad scale) live in device tables the captured kernels read, so LR schedules need no
re-capture; the mask noise comes from torch's graph-safe Philox generator state."""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, trainer, image_example: torch.Tensor, audio_example: torch.Tensor, warmup: int = 3,
                 capture_error_mode: str = "thread_local"):
        assert trainer.accum_iter == 1, "graph capture covers one full optimizer step (accum_iter == 1)"
        self.trainer = trainer
        self.image = torch.empty_like(image_example)
        self.audio = torch.empty_like(audio_example)
        self.image.copy_(image_example)
        self.audio.copy_(audio_example)
        opt = trainer.optimizer
        # eager warm-up on a side stream: instantiates kernels / attributes / TMA descriptors / allocator pools
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._one_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        opt._sync_hp()
        self.graph = torch.cuda.CUDAGraph()
        from .. import kernels as K
        n0 = K.launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.loss_image, self.loss_audio, self.grad_norm = self._one_step(sync_hp=False)
        self.launches_per_step = K.launch_count() - n0

    def _one_step(self, sync_hp: bool = True):
        tr = self.trainer
        li, la, _, _ = tr.model(self.image, self.audio)
        tr.backward(li + la)
        tr.optimizer.step(zero_grad=True, sync_hp=sync_hp)
        tr.accums = 0
        return li.detach(), la.detach(), tr.optimizer.grad_norm()

    def __call__(self, image: torch.Tensor, audio: torch.Tensor):
        """image / audio may live on the host (pinned): the copies run on the current stream before the replay."""
        self.image.copy_(image, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        self.trainer.optimizer._sync_hp()
        self.graph.replay()
        self.trainer.optimizer.n_steps += 1
        self.trainer.n_steps += 1
        return self.loss_image, self.loss_audio, self.grad_norm
