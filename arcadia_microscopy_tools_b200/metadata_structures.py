"""Acquisition metadata containers needed to build a ``MicroscopyImage``.

Pure metadata (no pixels): the field names follow the reference's
``metadata_structures.py:34-141`` so objects built for the reference can be rebuilt here
unchanged.  Vendor metadata *parsing* (ND2 / LIF) is out of scope for the B200 path.
"""

from __future__ import annotations

from dataclasses import dataclass, field, fields
from datetime import datetime
from enum import Flag, auto
from typing import Any

from .channels import Channel


class DimensionFlags(Flag):
    """Which acquisition dimensions an image has beyond the 2-D plane."""

    SPATIAL_2D = 0
    MULTICHANNEL = auto()
    Z_STACK = auto()
    TIMELAPSE = auto()
    SPECTRAL = auto()
    RGB = auto()
    MONTAGE = auto()

    @property
    def is_multichannel(self) -> bool:
        return DimensionFlags.MULTICHANNEL in self

    @property
    def is_zstack(self) -> bool:
        return DimensionFlags.Z_STACK in self

    @property
    def is_timelapse(self) -> bool:
        return DimensionFlags.TIMELAPSE in self

    @property
    def is_spectral(self) -> bool:
        return DimensionFlags.SPECTRAL in self

    @property
    def is_rgb(self) -> bool:
        return DimensionFlags.RGB in self

    @property
    def is_montage(self) -> bool:
        return DimensionFlags.MONTAGE in self


def _needs(dimension: DimensionFlags, default: Any = None) -> Any:
    """Dataclass field that becomes mandatory when ``dimension`` is present."""
    return field(default=default, metadata={"requires_dimension": dimension})


class _DimensionChecked:
    def validate(self, dimensions: DimensionFlags) -> None:
        for f in fields(self):  # type: ignore[arg-type]
            required = f.metadata.get("requires_dimension")
            if required and (dimensions & required) and getattr(self, f.name) is None:
                raise ValueError(f"{f.name} is required for {required.name}")


@dataclass
class NominalDimensions(_DimensionChecked):
    x_size_px: int
    y_size_px: int
    xy_step_um: float
    z_size_px: int | None = _needs(DimensionFlags.Z_STACK)
    z_step_um: float | None = _needs(DimensionFlags.Z_STACK)
    t_size_px: int | None = _needs(DimensionFlags.TIMELAPSE)
    t_step_ms: float | None = _needs(DimensionFlags.TIMELAPSE)
    w_size_px: int | None = _needs(DimensionFlags.SPECTRAL)
    w_step_nm: float | None = _needs(DimensionFlags.SPECTRAL)


@dataclass
class MeasuredDimensions(_DimensionChecked):
    x_values_um: Any = _needs(DimensionFlags.MONTAGE)
    y_values_um: Any = _needs(DimensionFlags.MONTAGE)
    z_values_um: Any = _needs(DimensionFlags.Z_STACK)
    t_values_ms: Any = _needs(DimensionFlags.TIMELAPSE)
    w_values_nm: Any = _needs(DimensionFlags.SPECTRAL)


@dataclass
class AcquisitionSettings(_DimensionChecked):
    exposure_time_s: float | None = None
    zoom: float | None = None
    binning: str | None = None
    pixel_dwell_time_us: float | None = None
    line_scan_speed_hz: float | None = None
    line_averaging: int | None = None
    line_accumulation: int | None = None
    frame_averaging: int | None = None
    frame_accumulation: int | None = None


@dataclass
class MicroscopeConfig:
    magnification: int
    numerical_aperture: float
    objective: str | None = None
    light_source: str | None = None
    power_mw: float | None = None


@dataclass
class ChannelMetadata:
    channel: Channel
    timestamp: datetime
    dimensions: DimensionFlags
    resolution: NominalDimensions
    measured: MeasuredDimensions
    acquisition: AcquisitionSettings
    optics: MicroscopeConfig

    def __post_init__(self) -> None:
        self.resolution.validate(self.dimensions)
        self.measured.validate(self.dimensions)

    @classmethod
    def minimal(cls, channel: Channel, height: int, width: int, dimensions: DimensionFlags = DimensionFlags.SPATIAL_2D,
                xy_step_um: float = 1.0, **resolution: Any) -> "ChannelMetadata":
        """Smallest valid record for a channel (synthetic data, raw ND2 frames)."""
        measured = MeasuredDimensions(
            z_values_um=[0.0] * int(resolution.get("z_size_px") or 0) if dimensions.is_zstack else None,
            t_values_ms=[0.0] * int(resolution.get("t_size_px") or 0) if dimensions.is_timelapse else None,
        )
        return cls(
            channel=channel,
            timestamp=datetime.fromtimestamp(0),
            dimensions=dimensions,
            resolution=NominalDimensions(x_size_px=width, y_size_px=height, xy_step_um=xy_step_um, **resolution),
            measured=measured,
            acquisition=AcquisitionSettings(),
            optics=MicroscopeConfig(magnification=1, numerical_aperture=1.0),
        )
