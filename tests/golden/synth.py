"""Re-export of pnp_ovss_b200.synthetic (kept so the golden generator and the tests can `import synth`)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from pnp_ovss_b200.synthetic import *  # noqa: F401,F403,E402
from pnp_ovss_b200.synthetic import CLS_ID, ENC_ID, PAD_ID, SEP_ID  # noqa: F401,E402
