"""Load tests/golden/*.npz (written by tests/golden/make_golden.py) as {case: {key: array}}."""
import os
from collections import defaultdict

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases(name):
    z = np.load(os.path.join(GOLDEN, name))
    cases = defaultdict(dict)
    for key in z.files:
        case, field = key.split("/", 1)
        cases[case][field] = z[key]
    return dict(cases)
