"""Names models/swin.py:5 imports; the Swin decoder is out of scope (decoder_arch: plain), bodies are stubs."""


def get_relative_position_index(win_h, win_w):
    raise NotImplementedError("swin decoder is out of scope")


def window_partition(x, window_size):
    raise NotImplementedError("swin decoder is out of scope")


def window_reverse(windows, window_size, H, W):
    raise NotImplementedError("swin decoder is out of scope")
