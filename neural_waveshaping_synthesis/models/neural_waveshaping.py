from neural_waveshaping_synthesis_b200.models.neural_waveshaping import *  # noqa
from neural_waveshaping_synthesis_b200.models.neural_waveshaping import ControlModule, NeuralWaveshaping  # noqa
