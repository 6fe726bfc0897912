"""Channel identity type.

Only the *type* matters on the hot path: channels key ``intensity_image_dict`` and their
lower-cased name suffixes the per-channel columns (ref: ``masks.py:326``).  Mirrors the frozen,
hashable ``Channel`` dataclass and the predefined channel table of the reference's
``channels.py:35-117``.  Colour science (``colour``) is not a dependency here: deriving a hex
colour from a wavelength is visualisation and stays out of scope, so ``from_wavelength``
imports it lazily and says so when it is absent.
"""

from __future__ import annotations

import re
from dataclasses import dataclass

_HEX_COLOR = re.compile(r"^#(?:[0-9a-fA-F]{3}){1,2}$")


@dataclass(frozen=True)
class Channel:
    """An imaging channel: a name, a display colour, and optional wavelengths in nm."""

    name: str
    color: str
    excitation_nm: float | None = None
    emission_nm: float | None = None

    def __post_init__(self) -> None:
        if _HEX_COLOR.match(self.color) is None:
            raise ValueError(f"color must be a hex code like '#FF0000', got '{self.color}'")
        for label, value in (("excitation_nm", self.excitation_nm), ("emission_nm", self.emission_nm)):
            if value is not None and value <= 0:
                raise ValueError(f"{label} must be positive")

    @classmethod
    def from_wavelength(cls, wavelength_nm: float, *, name: str | None = None, is_excitation: bool = True) -> "Channel":
        """Channel whose colour is derived from a visible wavelength (needs ``colour``)."""
        if not 360 <= wavelength_nm <= 780:
            raise ValueError(f"Wavelength must be in the visible range (360-780 nm), got {wavelength_nm} nm")
        try:
            import colour  # type: ignore
            import numpy as np
        except ImportError as exc:  # pragma: no cover - optional dependency
            raise ImportError("Channel.from_wavelength needs the optional 'colour-science' package") from exc
        rgb = np.clip(colour.XYZ_to_sRGB(colour.wavelength_to_XYZ(wavelength_nm)), 0, 1)
        r, g, b = (rgb * 255).astype(int)
        wl = round(wavelength_nm, 1)
        return cls(
            name=name or f"{wavelength_nm:.0f}nm",
            color=f"#{r:02X}{g:02X}{b:02X}",
            excitation_nm=wl if is_excitation else None,
            emission_nm=None if is_excitation else wl,
        )


def _table() -> dict[str, Channel]:
    rows = [
        ("BRIGHTFIELD", "#FFFFFF", None, None),
        ("DIC", "#FFFFFF", None, None),
        ("PHASE", "#DDDDDD", None, None),
        ("DAPI", "#0033FF", 405, 450),
        ("FITC", "#07FF00", 488, 512),
        ("TRITC", "#FFBF00", 561, 595),
        ("CY5", "#A30000", 640, 665),
        ("SRS", "#E63535", None, None),
        ("E-CARS", "#AB1299", None, None),
        ("F-CARS", "#AB1299", None, None),
        ("E-SHG", "#F29B4F", None, None),
        ("F-SHG", "#F29B4F", None, None),
    ]
    return {n: Channel(n, col, ex, em) for n, col, ex, em in rows}


CHANNELS: dict[str, Channel] = _table()
BRIGHTFIELD = CHANNELS["BRIGHTFIELD"]
DIC = CHANNELS["DIC"]
PHASE = CHANNELS["PHASE"]
DAPI = CHANNELS["DAPI"]
FITC = CHANNELS["FITC"]
TRITC = CHANNELS["TRITC"]
CY5 = CHANNELS["CY5"]
SRS = CHANNELS["SRS"]
E_CARS = CHANNELS["E-CARS"]
F_CARS = CHANNELS["F-CARS"]
E_SHG = CHANNELS["E-SHG"]
F_SHG = CHANNELS["F-SHG"]
