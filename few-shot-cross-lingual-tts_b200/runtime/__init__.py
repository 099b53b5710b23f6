from .dp import GradBuckets
from .optim import FusedAdam
from .step import TrainStep
from .fomaml import FirstOrderTaskStep
