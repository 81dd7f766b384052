from neural_waveshaping_synthesis_b200.models.modules.dynamic import *  # noqa
