from .dp import GradBuckets
from .optim import FusedAdam
from .step import TrainStep
from .fomaml import FirstOrderTaskStep
from .cache import TrainStepCache, pad_batch
