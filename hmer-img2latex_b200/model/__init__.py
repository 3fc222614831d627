from .encoder import CNNEncoder, ResNetEncoder, normalize_u8
from .decoder import LSTMDecoder, Attention
from .seq2seq import Seq2SeqModel

__all__ = ["CNNEncoder", "ResNetEncoder", "LSTMDecoder", "Attention", "Seq2SeqModel", "normalize_u8"]
