"""Drop-in import name: `import np_bnn as bn` resolves to the B200-native implementation of the MCMC hot
path (npbnn_b200).  Only the names on that path are provided; see DESIGN.md for what is out of scope."""
from npbnn_b200 import *  # noqa: F401,F403
from npbnn_b200 import __version__  # noqa: F401
