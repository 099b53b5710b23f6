from .fastspeech2m import FastSpeech2
from .loss import FastSpeech2Loss, FastSpeech2ADALoss
