"""Image container: pixels plus acquisition metadata, channel slicing, pipeline dispatch.

Drop-in for the reference's ``microscopy.py`` (``InstrumentMetadata`` :17-67, ``Metadata``
:70-88, ``MicroscopyImage`` :91-308): same attributes, same validation and messages, same
view semantics for ``get_channel_intensities``.  The container does no arithmetic; the work is
in ``Pipeline`` / ``operations`` (GPU).  ``from_nd2_path`` uses the package's own raw-frame
reader (pixels + channel order only; vendor metadata parsing is out of scope).
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from functools import cached_property
from pathlib import Path
from typing import Any

import numpy as np

from .channels import CHANNELS, Channel
from .exceptions import MetadataWarning
from .metadata_structures import ChannelMetadata, DimensionFlags
from .pipeline import Pipeline


@dataclass
class InstrumentMetadata:
    """Dimension sizes (ordered like the array axes) and one metadata record per channel."""

    sizes: dict[str, int]
    channel_metadata_list: list[ChannelMetadata]

    def __post_init__(self) -> None:
        missing = [k for k in ("X", "Y") if k not in self.sizes]
        if missing:
            raise ValueError(
                f"sizes must contain '{missing[0]}' dimension, got keys: {list(self.sizes.keys())}"
            )
        expected = self.sizes.get("C", 1)
        actual = len(self.channel_metadata_list)
        if actual != expected:
            raise ValueError(
                f"Number of channel metadata entries ({actual}) does not match "
                f"the channel dimension size ({expected}) in sizes"
            )

    @property
    def channel_axis(self) -> int | None:
        keys = list(self.sizes.keys())
        return keys.index("C") if "C" in keys else None

    @cached_property
    def dimensions(self) -> DimensionFlags:
        flags = DimensionFlags(0)
        for record in self.channel_metadata_list:
            flags |= record.dimensions
        if len(self.channel_metadata_list) > 1:
            flags |= DimensionFlags.MULTICHANNEL
        return flags


@dataclass
class Metadata:
    """Instrument metadata plus optional free-form sample metadata."""

    instrument: InstrumentMetadata
    sample: dict[str, Any] | None = None

    def __repr__(self) -> str:
        names = [record.channel.name for record in self.instrument.channel_metadata_list]
        tail = f", sample={self.sample}" if self.sample else ""
        return f"<Metadata sizes={self.instrument.sizes}, channels={names}{tail}>"


@dataclass
class MicroscopyImage:
    """uint16 intensities of every channel together with their metadata."""

    intensities: np.ndarray
    metadata: Metadata

    def __post_init__(self) -> None:
        expected_shape = tuple(self.metadata.instrument.sizes.values())
        if self.intensities.shape != expected_shape:
            raise ValueError(
                f"Intensities shape {self.intensities.shape} does not match "
                f"metadata sizes {self.metadata.instrument.sizes} "
                f"(expected shape {expected_shape})"
            )
        if self.intensities.dtype != np.uint16:
            warnings.warn(
                f"Expected uint16 intensities, got {self.intensities.dtype}. "
                f"Some operations may behave unexpectedly.",
                MetadataWarning,
                stacklevel=2,
            )

    def __repr__(self) -> str:
        flat = self.intensities.flat
        if self.intensities.size <= 10:
            shown = f"intensities={list(flat)}"
        else:
            head = ", ".join(str(v) for v in flat[:3].tolist())
            tail = ", ".join(str(v) for v in flat[-3:].tolist())
            shown = f"intensities=[{head}, ..., {tail}]"
        names = [channel.name for channel in self.channels]
        return f"<MicroscopyImage sizes={self.sizes}, channels={names}, {shown}, dtype={self.intensities.dtype}>"

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_arrays(
        cls,
        intensities: np.ndarray,
        channels: list[Channel],
        axes: str,
        sample_metadata: dict[str, Any] | None = None,
        xy_step_um: float = 1.0,
    ) -> "MicroscopyImage":
        """Wrap a pixel block whose axes are named by ``axes`` (e.g. ``"CYX"``, ``"TCYX"``)."""
        if len(axes) != intensities.ndim:
            raise ValueError(f"axes '{axes}' does not describe a {intensities.ndim}-D array")
        sizes = {ax: int(n) for ax, n in zip(axes, intensities.shape)}
        dims = DimensionFlags(0)
        extra: dict[str, Any] = {}
        if "T" in sizes:
            dims |= DimensionFlags.TIMELAPSE
            extra.update(t_size_px=sizes["T"], t_step_ms=0.0)
        if "Z" in sizes:
            dims |= DimensionFlags.Z_STACK
            extra.update(z_size_px=sizes["Z"], z_step_um=1.0)
        records = []
        for channel in channels:
            records.append(ChannelMetadata.minimal(channel, sizes["Y"], sizes["X"], dims, xy_step_um, **extra))
        return cls(intensities, Metadata(InstrumentMetadata(sizes, records), sample_metadata))

    @classmethod
    def from_nd2_path(
        cls,
        nd2_path: Path,
        channels: list[Channel] | None = None,
        sample_metadata: dict[str, Any] | None = None,
    ) -> "MicroscopyImage":
        """Load the raw frames of an uncompressed ND2 file (ref: ``microscopy.py:154-176``).

        ``channels`` names the components in file order; it is required for multi-component
        files because optical-configuration parsing is out of scope here.
        """
        from .nd2_raw import read_nd2_frames

        frames = read_nd2_frames(nd2_path)  # (frames, C, Y, X)
        n_frames, n_comp = frames.shape[:2]
        if channels is None:
            if n_comp != 1:
                raise ValueError("channels must be given for multi-component ND2 files")
            channels = [CHANNELS["BRIGHTFIELD"]]
        if len(channels) != n_comp:
            raise ValueError(f"{len(channels)} channels given for {n_comp} components")
        axes = "CYX"
        data = frames
        if n_frames > 1:
            axes = "T" + axes
        else:
            data = data[0]
        if n_comp == 1:
            data = data[:, 0] if n_frames > 1 else data[0]
            axes = axes.replace("C", "")
        return cls.from_arrays(np.ascontiguousarray(data), channels, axes, sample_metadata)

    @classmethod
    def from_lif_path(
        cls,
        lif_path: Path,
        image_name: str,
        channels: list[Channel] | None = None,
        sample_metadata: dict[str, Any] | None = None,
    ) -> "MicroscopyImage":
        """Load the pixel block of one image of a Leica LIF file (ref: ``microscopy.py:178-202``).

        ``channels`` names the channels in file order; it is required for multi-channel images because
        detector / laser parsing (``leica.py``) is out of scope here.  Supported axes: T, Z, C, Y, X in the
        order the file stores them.
        """
        from .lif_raw import read_lif_image

        data, sizes = read_lif_image(lif_path, image_name)
        unsupported = [ax for ax in sizes if ax not in "TZCYX"]
        if unsupported:
            raise ValueError(f"image '{image_name}' has axes {list(sizes)}; only T, Z, C, Y, X are supported")
        n_channels = sizes.get("C", 1)
        if channels is None:
            if n_channels != 1:
                raise ValueError("channels must be given for multi-channel LIF images")
            channels = [CHANNELS["BRIGHTFIELD"]]
        if len(channels) != n_channels:
            raise ValueError(f"{len(channels)} channels given for {n_channels} channels in the file")
        return cls.from_arrays(data, channels, "".join(sizes), sample_metadata)

    # ------------------------------------------------------------------ views
    @property
    def shape(self) -> tuple[int, ...]:
        return self.intensities.shape

    @property
    def sizes(self) -> dict[str, int]:
        return self.metadata.instrument.sizes

    @property
    def dimensions(self) -> DimensionFlags:
        return self.metadata.instrument.dimensions

    @property
    def channels(self) -> list[Channel]:
        return [record.channel for record in self.metadata.instrument.channel_metadata_list]

    @property
    def channel_axis(self) -> int | None:
        return self.metadata.instrument.channel_axis

    @property
    def num_channels(self) -> int:
        return len(self.metadata.instrument.channel_metadata_list)

    @staticmethod
    def _resolve_channel_name(channel: str | Channel) -> str:
        return channel if isinstance(channel, str) else channel.name

    def get_channel_intensities(self, channel: str | Channel) -> np.ndarray:
        """All data of one channel, as a view (ref: ``microscopy.py:241-282``)."""
        name = self._resolve_channel_name(channel)
        names = [ch.name for ch in self.channels]
        if name not in names:
            raise ValueError(f"Channel '{name}' not found in image. Available channels: {names}")
        if self.num_channels == 1:
            return self.intensities
        axis = self.channel_axis
        if axis is None:
            raise ValueError("Channel axis not found in metadata")
        index: list[slice | int] = [slice(None)] * self.intensities.ndim
        index[axis] = names.index(name)
        return self.intensities[tuple(index)]

    def apply_pipeline(self, pipeline: Pipeline, channel: str | Channel):
        """Run ``pipeline`` on one channel's data (ref: ``microscopy.py:284-308``)."""
        return pipeline(self.get_channel_intensities(channel))
