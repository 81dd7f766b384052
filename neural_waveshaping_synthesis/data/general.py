from neural_waveshaping_synthesis_b200.data.general import GeneralDataset  # noqa
