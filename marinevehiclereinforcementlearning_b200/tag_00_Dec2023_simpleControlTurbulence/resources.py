"""``headingError`` of tag_00_Dec2023_simpleControlTurbulence/resources.py:26-46
(same function as the current ``resources.angleError``); runs ``mvrl_angle_error``."""
from ..resources import angleError as headingError  # noqa: F401
from ..resources import evaluate_agent  # noqa: F401  (tag_00.../resources.py:49-101)

orientation = "right_up_anticlockwise"  # tag_00.../resources.py:23
