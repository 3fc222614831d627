"""hmer-img2latex_b200 -- B200-native (sm_100a) implementation of the batched inference
path of Jeremy-Cleland/hmer-img2latex behind the reference's own model / decoder API.
Hand-written CUDA kernels behind a C-ABI (include/i2l_b200.h); PyTorch only provides
device memory, streams and torch.distributed.  CUDA only: no CPU fallback."""
from . import _native
from .model import Attention, CNNEncoder, LSTMDecoder, ResNetEncoder, Seq2SeqModel, normalize_u8
from . import metrics, preprocess
from .predictor import Predictor
from .preprocess import ResizeWithAspectRatio, load_images
from .tokenizer import LaTeXTokenizer

__all__ = ["Attention", "CNNEncoder", "LSTMDecoder", "ResNetEncoder", "Seq2SeqModel", "Predictor",
           "LaTeXTokenizer", "normalize_u8", "ResizeWithAspectRatio", "load_images", "preprocess", "metrics", "_native"]
__version__ = "0.1.0"
