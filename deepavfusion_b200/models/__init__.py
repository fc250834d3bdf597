from .deepavfusion import DeepAVFusion  # noqa: F401
from .avmae import AVMAE  # noqa: F401
from .classifier import AVClassifier  # noqa: F401
