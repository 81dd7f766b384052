"""`import gin` for the drop-in callers (scripts/time_forward_pass.py:4).

gin-config is not installable here (no network, not in /opt/wheelhouse), so the
repo ships a minimal stand-in: neural_waveshaping_synthesis_b200/compat/gin_shim.py.
This directory doubles as the config-file tree the reference scripts are pointed
at (``--gin-file gin/models/newt.gin``).
"""
from neural_waveshaping_synthesis_b200.compat.gin_shim import *  # noqa: F401,F403
from neural_waveshaping_synthesis_b200.compat.gin_shim import (  # noqa: F401
    REQUIRED, config_scope, configurable, external_configurable, parse_config_file,
)
