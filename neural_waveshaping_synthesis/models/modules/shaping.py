from neural_waveshaping_synthesis_b200.models.modules.shaping import *  # noqa
from neural_waveshaping_synthesis_b200.models.modules.shaping import FastNEWT, NEWT, Reverb, Sine, TrainableNonlinearity  # noqa
