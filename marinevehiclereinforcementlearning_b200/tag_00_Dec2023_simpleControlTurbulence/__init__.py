"""Drop-in for the reference's frozen legacy image
``tag_00_Dec2023_simpleControlTurbulence/``: ``verySimpleAuv`` (AuvEnv,
PDController, make_env), ``verySimpleAuv_cyl`` (AuvEnvCyl), ``flowGenerator`` (ReconstructedFlow) and the
``headingError`` helper of its ``resources.py``.  The SB3 training drivers,
plotting and analysis scripts of that directory are out of scope."""
from . import flowGenerator, resources, verySimpleAuv, verySimpleAuv_cyl  # noqa: F401
