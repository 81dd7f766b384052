from neural_waveshaping_synthesis_b200.data.utils.upsampling import (  # noqa
    cubic_spline_interpolation, interp_frames_batch, linear_interpolation, overlap_add_upsample)
