"""Warning categories (same names and base class as the reference's ``exceptions.py:1-6``)."""


class MetadataWarning(UserWarning):
    """Image metadata was missing or ambiguous and a fallback was used."""


class SegmentationWarning(UserWarning):
    """A segmentation step produced a degraded or missing result."""
