"""CPU oracle for the preprocess -> threshold/label -> quantify hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU baseline.
The product package (``arcadia_microscopy_tools_b200``) never imports this package and
fails loudly when its CUDA library is missing.

What it restates
----------------
The reference (``/root/reference``, arcadia-microscopy-tools v0.4.1) is pure Python and
delegates every pixel loop to un-vendored third-party libraries pinned in its
``uv.lock``: scikit-image 0.25.2, scipy 1.16.1, numpy 2.4.1.  scikit-image cannot be
installed in this environment (no wheel, no network), so each scikit-image function on
the path is restated here from its published algorithm on top of the scipy / numpy
routines it dispatches to (which ARE importable here: scipy 1.18.1, numpy 2.3.5), and
the reference's own call sequences (``operations.py``, ``masks.py``) are restated on
top of those.  Every function cites the reference file:line it follows.

Pinning status
--------------
* Gaussian / DoG, bool labelling, percentiles, float histograms: pinned bit-for-bit
  against the real third-party code the reference runs (``scipy.ndimage.gaussian_filter``,
  ``scipy.ndimage.label``, ``np.percentile``, ``np.histogram``) in ``tests/test_oracle.py``.
* Raw-data golden checks (sha256, per-channel min/max/sum) and the provisional
  known answers of SURVEY.md section 8c for ``example-multichannel.nd2`` are checked in
  ``tests/test_golden.py`` against ``tests/golden/``.
* Independent implementations of the same definitions in OpenCV 4.13 pin further legs
  (``tests/test_oracle_opencv.py``): integer-image Otsu, 8-connected labelling with area /
  bbox / centroid, second-order central moments, the Gaussian's definition (nearest and
  reflect borders), the box mean / std behind niblack and sauvola, clear_border (bool and
  integer-fragment semantics) and relabel_sequential; ``outlines.extract_outlines_cellpose``
  is a loop around the REAL ``cv2.findContours`` (``tests/test_outlines.py``).
* What remains scikit-image's *own* Python with no importable implementation here
  (isodata / yen / li / minimum / triangle scans, the float-image Otsu path, the
  threshold_local / niblack / sauvola formulas, find_contours' case table and segment
  joining, perimeter weights, 3-D regionprops formulas) is restated from the published
  source: **parity unpinned** for those legs beyond the reference's coarse assertions (disc
  areas, centroids within 2 px, circularity range, outline contract), scikit-image's own
  ``find_contours`` docstring example and the committed golden vectors of the reference
  fixture (``tests/golden/``), which ``tests/test_oracle.py``, ``tests/test_outlines.py`` and
  ``tests/test_golden*.py`` re-check.
"""

from . import exposure, filters, labeling, outlines, percentile, regionprops, threshold  # noqa: F401
from .ops import (  # noqa: F401
    apply_threshold,
    cell_properties,
    crop_to_center,
    process_mask,
    rescale_by_percentile,
    subtract_background_dog,
)
