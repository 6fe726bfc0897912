"""NumPy array aliases used across the package (mirrors the names of the reference's
``typing.py:5-13`` so annotations written against it keep working)."""

import numpy as np
from numpy.typing import NDArray

BoolArray = NDArray[np.bool_]
UByteArray = NDArray[np.uint8]
UInt16Array = NDArray[np.uint16]
Int64Array = NDArray[np.int64]
Float32Array = NDArray[np.float32]
Float64Array = NDArray[np.float64]

ScalarArray = BoolArray | UByteArray | UInt16Array | Int64Array | Float32Array | Float64Array
