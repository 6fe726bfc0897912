"""Oracle: cell outlines (ref: ``masks.py:68-115``, ``masks.py:229-245``).  TEST INFRASTRUCTURE ONLY.

* ``extract_outlines_cellpose``: ``cellpose.utils.outlines_list(masks, multiprocessing=False)``
  (cellpose 4.0.8, ``outlines_list_single``) is a loop around the REAL ``cv2.findContours`` (importable
  here: OpenCV 4.13; the reference pins opencv-python-headless 4.11) — pinned to the real library.
* ``extract_outlines_skimage``: ``skimage.measure.find_contours(crop, 0.5)`` restated from the published
  source (``_find_contours_cy._get_contour_segments`` + ``_find_contours._assemble_contours``); scikit-image
  is not installable here, so this leg is **parity unpinned** beyond the reference's own assertions
  (``tests/test_masks.py:86-149``: closed, one per cell, image coordinates, centroid within 2 px).
"""

from __future__ import annotations

from collections import deque

import numpy as np
from scipy import ndimage as ndi


def outlines_list_single(masks: np.ndarray) -> list[np.ndarray]:
    """cellpose ``utils.outlines_list_single``: per label the longest external OpenCV contour as (x, y)
    integer points; fewer than 5 points -> ``np.zeros((0, 2))``."""
    import cv2

    outpix = []
    for n in np.unique(masks)[1:]:
        mn = masks == n
        if mn.sum() > 0:
            contours = cv2.findContours(mn.astype(np.uint8), mode=cv2.RETR_EXTERNAL, method=cv2.CHAIN_APPROX_NONE)
            contours = contours[-2]
            cmax = np.argmax([c.shape[0] for c in contours])
            pix = contours[cmax].astype(int).squeeze()
            if len(pix) > 4:
                outpix.append(pix)
            else:
                outpix.append(np.zeros((0, 2)))
    return outpix


def extract_outlines_cellpose(label_image: np.ndarray) -> list[np.ndarray]:
    """ref: masks.py:68-79 — cellpose's (x, y) points flipped to (y, x)."""
    outlines = outlines_list_single(label_image)
    return [outline[:, [1, 0]] if len(outline) > 0 else outline for outline in outlines]


def _get_fraction(from_value: float, to_value: float, level: float) -> float:
    if to_value == from_value:
        return 0.0
    return (level - from_value) / (to_value - from_value)


def _get_contour_segments(array: np.ndarray, level: float, vertex_connect_high: bool = False):
    segments = []
    for r0 in range(array.shape[0] - 1):
        for c0 in range(array.shape[1] - 1):
            r1, c1 = r0 + 1, c0 + 1
            ul, ur, ll, lr = array[r0, c0], array[r0, c1], array[r1, c0], array[r1, c1]
            square_case = 0
            if ul > level:
                square_case += 1
            if ur > level:
                square_case += 2
            if ll > level:
                square_case += 4
            if lr > level:
                square_case += 8
            if square_case in (0, 15):
                continue
            top = (r0, c0 + _get_fraction(ul, ur, level))
            bottom = (r1, c0 + _get_fraction(ll, lr, level))
            left = (r0 + _get_fraction(ul, ll, level), c0)
            right = (r0 + _get_fraction(ur, lr, level), c1)
            if square_case == 1:
                segments.append((top, left))
            elif square_case == 2:
                segments.append((right, top))
            elif square_case == 3:
                segments.append((right, left))
            elif square_case == 4:
                segments.append((left, bottom))
            elif square_case == 5:
                segments.append((top, bottom))
            elif square_case == 6:
                if vertex_connect_high:
                    segments.append((left, top))
                    segments.append((right, bottom))
                else:
                    segments.append((right, top))
                    segments.append((left, bottom))
            elif square_case == 7:
                segments.append((right, bottom))
            elif square_case == 8:
                segments.append((bottom, right))
            elif square_case == 9:
                if vertex_connect_high:
                    segments.append((top, right))
                    segments.append((bottom, left))
                else:
                    segments.append((top, left))
                    segments.append((bottom, right))
            elif square_case == 10:
                segments.append((bottom, top))
            elif square_case == 11:
                segments.append((bottom, left))
            elif square_case == 12:
                segments.append((left, right))
            elif square_case == 13:
                segments.append((top, right))
            elif square_case == 14:
                segments.append((left, top))
    return segments


def _assemble_contours(segments):
    current_index = 0
    contours = {}
    starts = {}
    ends = {}
    for from_point, to_point in segments:
        if from_point == to_point:
            continue
        tail, tail_num = starts.pop(to_point, (None, None))
        head, head_num = ends.pop(from_point, (None, None))
        if tail is not None and head is not None:
            if tail is head:
                head.append(to_point)
            elif tail_num > head_num:
                head.extend(tail)
                contours.pop(tail_num, None)
                starts[head[0]] = (head, head_num)
                ends[head[-1]] = (head, head_num)
            else:
                tail.extendleft(reversed(head))
                starts.pop(head[0], None)
                contours.pop(head_num, None)
                starts[tail[0]] = (tail, tail_num)
                ends[tail[-1]] = (tail, tail_num)
        elif tail is None and head is None:
            new_contour = deque((from_point, to_point))
            contours[current_index] = new_contour
            starts[from_point] = (new_contour, current_index)
            ends[to_point] = (new_contour, current_index)
            current_index += 1
        elif head is None:
            tail.appendleft(from_point)
            starts[from_point] = (tail, tail_num)
        else:
            head.append(to_point)
            ends[to_point] = (head, head_num)
    return [np.array(contour) for _, contour in sorted(contours.items())]


def find_contours(image: np.ndarray, level: float = 0.5) -> list[np.ndarray]:
    """``skimage.measure.find_contours(image, level)`` with the defaults fully_connected='low',
    positive_orientation='low', mask=None."""
    image = np.asarray(image, dtype=np.float64)
    return _assemble_contours(_get_contour_segments(image, float(level), False))


def extract_outlines_skimage(label_image: np.ndarray) -> list[np.ndarray]:
    """ref: masks.py:82-115 — per region (ascending label) ``find_contours`` on the 1-px padded crop
    ``label_image == label``, longest contour, shifted back to image coordinates."""
    h, w = label_image.shape
    outlines = []
    for index, slices in enumerate(ndi.find_objects(label_image)):
        if slices is None:
            continue
        label = index + 1
        minr, minc, maxr, maxc = slices[0].start, slices[1].start, slices[0].stop, slices[1].stop
        minr_p, minc_p = max(minr - 1, 0), max(minc - 1, 0)
        maxr_p, maxc_p = min(maxr + 1, h), min(maxc + 1, w)
        crop = (label_image[minr_p:maxr_p, minc_p:maxc_p] == label).astype(np.uint8)
        contours = find_contours(crop, level=0.5)
        if contours:
            main_contour = max(contours, key=len)
            main_contour = main_contour + np.array([minr_p, minc_p])
            outlines.append(main_contour)
        else:
            outlines.append(np.array([]).reshape(0, 2))
    return outlines
