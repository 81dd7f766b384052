from neural_waveshaping_synthesis_b200.data.utils.loudness_extraction import (  # noqa
    compute_power_spectrogram, extract_perceptual_loudness, extract_rms, perform_perceptual_weighting,
    perceptual_loudness_batch, rms_batch)
