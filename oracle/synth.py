"""Seeded synthetic checkpoints and inputs: alias of ``megatts2_hierspeechpp_b200.synthetic`` (the generator of
random-init weights / synthetic inputs is a utility of the package -- bench.py's B200 arm uses it -- and does not
import anything from ``oracle/``); kept here so the oracle-side tests read ``oracle.synth`` as before."""
import sys

from megatts2_hierspeechpp_b200 import synthetic as _synthetic

sys.modules[__name__] = _synthetic
