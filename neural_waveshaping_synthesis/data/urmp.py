from neural_waveshaping_synthesis_b200.data.urmp import URMPDataset  # noqa
