"""Operator / pipeline API: the drop-in boundary of the preprocessing path.

Same contract as the reference's ``pipeline.py:11-173`` (``ImageOperation`` = immutable
``(func, args, kwargs)`` callable; ``Pipeline`` = sequential composition with ``copy``,
``preserve_dtype``, ``parallel``, ``max_workers``), same error messages.  What changes is where
the work runs when every operation is one of this package's device operations
(``operations.*``): the input is uploaded once, the whole chain runs on the GPU, and the
result comes back once; with ``parallel=True`` the first axis becomes the batch axis of ONE set
of kernel launches instead of a ``ThreadPoolExecutor`` map (ref: ``pipeline.py:139-149``) —
slices are still processed independently, so results are identical.  Foreign callables keep the
reference's host behaviour (thread pool included).
"""

from __future__ import annotations

import warnings
from collections.abc import Callable, Sequence
from concurrent.futures import ThreadPoolExecutor
from typing import Any

import numpy as np


class ImageOperation:
    """A function plus the extra arguments to call it with on an intensity array."""

    __slots__ = ("func", "args", "kwargs")

    def __init__(self, func: Callable[..., Any], *args: object, **kwargs: object) -> None:
        for name, value in (("func", func), ("args", args), ("kwargs", kwargs)):
            object.__setattr__(self, name, value)

    def __setattr__(self, name: str, value: object) -> None:
        raise AttributeError("ImageOperation instances are immutable")

    def __delattr__(self, name: str) -> None:
        raise AttributeError("ImageOperation instances are immutable")

    def __call__(self, intensities):
        return self.func(intensities, *self.args, **self.kwargs)

    def _key(self) -> tuple:
        return (self.func, self.args, tuple(sorted(self.kwargs.items())))

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, ImageOperation):
            return NotImplemented
        return self.func == other.func and self.args == other.args and self.kwargs == other.kwargs

    def __hash__(self) -> int:
        return hash(self._key())

    def __repr__(self) -> str:
        parts = [repr(a) for a in self.args] + [f"{k}={v!r}" for k, v in self.kwargs.items()]
        return f"{getattr(self.func, '__name__', repr(self.func))}({', '.join(parts)})"

    @property
    def runs_on_device(self) -> bool:
        """True when ``func`` is one of this package's GPU operations."""
        return bool(getattr(self.func, "__amt_device_op__", False))


class Pipeline:
    """Apply image operations in order; optionally treat the first axis as independent slices."""

    def __init__(
        self,
        operations: Sequence[Callable[..., Any]],
        copy: bool = False,
        preserve_dtype: bool = False,
        parallel: bool = False,
        max_workers: int | None = None,
    ) -> None:
        self.operations = list(operations) if isinstance(operations, tuple) else operations
        self.copy = copy
        self.preserve_dtype = preserve_dtype
        self.parallel = parallel
        self.max_workers = max_workers
        if not self.operations:
            raise ValueError("Pipeline must have at least one operation")
        if not all(callable(op) for op in self.operations):
            raise TypeError("All operations must be callable (wrap functions with ImageOperation)")
        if max_workers is not None and max_workers < 1:
            raise ValueError(f"max_workers must be at least 1, got {max_workers}")
        if parallel and copy:
            warnings.warn(
                "copy=True has no effect when parallel=True. Parallel mode always produces a new output array.",
                UserWarning,
                stacklevel=2,
            )

    # ------------------------------------------------------------------ helpers
    def _device_chain(self) -> bool:
        return all(isinstance(op, ImageOperation) and op.runs_on_device for op in self.operations)

    def _apply_operations(self, intensities):
        out = intensities.copy() if self.copy else intensities
        for operation in self.operations:
            out = operation(out)
        return out

    def _apply_on_device(self, intensities: np.ndarray, batched: bool):
        from . import _gpu

        # the first operation uploads the NumPy array itself (its own dtype rules: uint8 keeps its 1/255 scale, wide
        # integers and bool are normalised); every intermediate then stays on the device
        out: Any = intensities
        _gpu._tls.keep_on_device = True
        try:
            for operation in self.operations:
                out = operation.func(out, *operation.args, **operation.kwargs, _batched=batched)
        finally:
            _gpu._tls.keep_on_device = False
        return _gpu.to_host(out) if _gpu.is_device_array(out) else out

    # ------------------------------------------------------------------ call
    def __call__(self, intensities):
        on_device = isinstance(intensities, np.ndarray) and intensities.size > 0 and self._device_chain()
        if self.parallel:
            if intensities.ndim < 3:
                raise ValueError(
                    f"Parallel mode requires at least 3D input (got {intensities.ndim}D). "
                    "The first axis is used to distribute work across threads."
                )
            if on_device:
                stacked = self._apply_on_device(intensities, batched=True)
            else:
                with ThreadPoolExecutor(max_workers=self.max_workers) as pool:
                    stacked = np.array(list(pool.map(self._apply_operations, intensities)))
            if self.preserve_dtype:
                return np.array(stacked, dtype=intensities.dtype)
            return stacked

        result = self._apply_on_device(intensities, batched=False) if on_device else self._apply_operations(intensities)
        if self.preserve_dtype and result.dtype != intensities.dtype:
            return result.astype(intensities.dtype)
        return result

    # ------------------------------------------------------------------ dunder
    def __len__(self) -> int:
        return len(self.operations)

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Pipeline):
            return NotImplemented
        mine = (self.operations, self.copy, self.preserve_dtype, self.parallel, self.max_workers)
        theirs = (other.operations, other.copy, other.preserve_dtype, other.parallel, other.max_workers)
        return mine == theirs

    def __repr__(self) -> str:
        flags = [
            text
            for text, on in (
                ("copy=True", self.copy),
                ("preserve_dtype=True", self.preserve_dtype),
                ("parallel=True", self.parallel),
                (f"max_workers={self.max_workers}", self.max_workers is not None),
            )
            if on
        ]
        ops = ", ".join(repr(op) for op in self.operations)
        return f"Pipeline([{ops}]{''.join(', ' + f for f in flags)})"
