"""helpers.py -- tile partitioner of the reference (helpers.py:5-17), kept for API parity.

The reference uses `divisor` to process rasters tile by tile on one GPU when they do not fit in
memory; its partitioned mode is broken for flow distance / slope.  The B200 entry
points accept the `division_*` keyword arguments and ignore them for the result: every call returns
what the reference returns with division_* = 0.
"""
import math

import numpy as np


def divisor(row_length, column_length, row_division, column_division):
    """Interior split indices floor((i+1)*L/(d+1)) -- helpers.py:5-17."""
    boundary_row = np.array([math.floor((i + 1) * row_length / (row_division + 1)) for i in range(row_division)], dtype=int)
    boundary_column = np.array(
        [math.floor((i + 1) * column_length / (column_division + 1)) for i in range(column_division)], dtype=int)
    return boundary_row, boundary_column
