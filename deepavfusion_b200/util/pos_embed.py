"""Fixed 2-D sin-cos positional embeddings (mirror of reference util/pos_embed.py:42-90).

Host-side, float32 numpy, computed once at construction; values are bit-identical to the
reference's (checked in oracle/make_golden.py), including its quirk that ``np.meshgrid(grid_w,
grid_h)`` puts the w coordinate first and the first half of the channels encodes it.
"""
import numpy as np


def _one_axis(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    assert embed_dim % 2 == 0
    freq = np.arange(embed_dim // 2, dtype=np.float32)
    freq /= embed_dim / 2.0
    freq = 1.0 / 10000 ** freq
    ang = np.einsum("m,d->md", pos.reshape(-1), freq)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def get_2d_sincos_pos_embed(embed_dim: int, grid_size, cls_token: bool = False) -> np.ndarray:
    """Returns [gH*gW, embed_dim] (or with a leading zero row when cls_token)."""
    if isinstance(grid_size, int):
        grid_size = (grid_size, grid_size)
    gh, gw = grid_size
    ys = np.arange(gh, dtype=np.float32)
    xs = np.arange(gw, dtype=np.float32)
    grid = np.stack(np.meshgrid(xs, ys), axis=0).reshape(2, 1, gh, gw)
    assert embed_dim % 2 == 0
    emb = np.concatenate([_one_axis(embed_dim // 2, grid[0]), _one_axis(embed_dim // 2, grid[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb
