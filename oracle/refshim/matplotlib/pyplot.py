"""pyplot stub: swallows every call; subplots() returns (fig, ndarray-of-axes)."""
import numpy as _np
from . import _Sink


def subplots(nrows=1, ncols=1, *_a, squeeze=True, **_k):
    fig = _Sink()
    if nrows == 1 and ncols == 1:
        return fig, _Sink()
    axes = _np.empty((nrows, ncols), dtype=object)
    for i in range(nrows):
        for j in range(ncols):
            axes[i, j] = _Sink()
    if squeeze:
        axes = axes.squeeze()
    return fig, axes


def __getattr__(_name):
    return _Sink()
