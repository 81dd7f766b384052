"""Import-path compatibility: the reference scripts import `neural_waveshaping_synthesis.*`
(scripts/time_forward_pass.py:10-11).  Every module here re-exports the B200 implementation from
neural_waveshaping_synthesis_b200 under the reference's module path."""
