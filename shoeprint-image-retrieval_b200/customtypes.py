"""Array and literal aliases shared by the package (reference: ``customtypes.py:7-16``)."""

from typing import Any, Literal

import numpy as np
import numpy.typing as npt

# one feature-map channel / one grayscale correlation surface
ImageArrayType = npt.NDArray[np.floating[Any]]
# a stack of channels, [C, h, w]
FeatureMapsArrayType = npt.NDArray[np.floating[Any]]

DatasetTypeType = Literal["FID-300", "Impress", "WVU2019"]
