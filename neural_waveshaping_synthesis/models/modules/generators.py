from neural_waveshaping_synthesis_b200.models.modules.generators import *  # noqa
