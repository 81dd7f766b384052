from neural_waveshaping_synthesis_b200.utils import make_dir_if_not_exists, seed_all  # noqa
