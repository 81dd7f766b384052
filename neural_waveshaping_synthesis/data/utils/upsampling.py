from neural_waveshaping_synthesis_b200.data.utils.upsampling import (  # noqa
    get_padded_length, get_source_target_axes, linear_interpolation)
