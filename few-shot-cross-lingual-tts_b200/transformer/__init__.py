from .Models import Encoder, Decoder, Encoder2
from .Layers import PostNet
