#!/usr/bin/env python
"""(Re)generate marinevehiclereinforcementlearning_b200/csrc/rov6_default_consts.h from the built library and rebuild:
    python tools/gen_default_consts.py
Same as `python -c "import __graft_entry__ as g; g.build()"` minus the oracle build."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marinevehiclereinforcementlearning_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    _lib.build(verbose=False)
    print("rov6_default_consts.h is up to date; library:", _lib.LIB_PATH)
