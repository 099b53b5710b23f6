from .dp import GradBuckets
from .step import TrainStep
