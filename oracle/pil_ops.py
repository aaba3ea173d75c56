"""TEST INFRASTRUCTURE (see oracle/__init__.py) - numpy restatement of the Pillow operations the reference's input
pipeline applies to every frame.

Reference call sites: ``data/image_pair_dataloader.py:72-165`` (rotate / resize / crop / flip / random filter, /255),
``data/keypoint_dataloader.py:58-82`` (resize + centre crop), ``utils/data.py:8-35`` (``apply_random_filter``),
``utils/data.py:38-59`` (``center_crop``), ``map_fn`` (``x * 2 - 1``, image_pair_dataloader.py:63-69).

The arithmetic lives in the third-party dependency **Pillow**, pinned by the reference at 6.2.0 (requirements.txt:10) and
not vendored; this container has Pillow 12.2.0.  What follows restates Pillow's C kernels (Geometry.c nearest-neighbour
affine / scale transforms in 16.16 fixed point, Filter.c 3x3 / 5x5 float kernels, Blend.c, Convert.c rgb->L) with the
6.2.0 *default arguments made explicit* (``resize`` and ``rotate`` resample = NEAREST; 7.0 changed ``resize``'s default
to BICUBIC).  ``tests/test_pil_ops_oracle.py`` pins every function here bit-for-bit against the installed Pillow called
with those arguments; C-level differences between 6.2.0 and 12.2.0, if any, cannot be checked here (no network).
"""
import math

import numpy as np

IMAGE_SIZE = 128

# ImageFilter built-ins used by utils/data.py:11-22: (size, scale, kernel)
KERNELS = {
    0: (3, 6, (0, -1, 0, -1, 10, -1, 0, -1, 0)),                      # DETAIL
    1: (3, 2, (-1, -1, -1, -1, 10, -1, -1, -1, -1)),                  # EDGE_ENHANCE
    2: (3, 13, (1, 1, 1, 1, 5, 1, 1, 1, 1)),                          # SMOOTH
    3: (5, 100, (1, 1, 1, 1, 1, 1, 5, 5, 5, 1, 1, 5, 44, 5, 1, 1, 5, 5, 5, 1, 1, 1, 1, 1, 1)),   # SMOOTH_MORE
    4: (3, 1, (-1, -1, -1, -1, 9, -1, -1, -1, -1)),                   # EDGE_ENHANCE_MORE
    5: (5, 16, (1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 1)),     # BLUR
}
# r_id 6..9 of apply_random_filter: ImageEnhance.{Sharpness, Brightness, Color, Contrast}
SHARPNESS, BRIGHTNESS, COLOR, CONTRAST = 6, 7, 8, 9


def _fix(v):
    """Geometry.c FIX(): double -> 16.16 fixed point, round half up via floor(v * 65536 + 0.5)."""
    return int(math.floor(v * 65536.0 + 0.5))


def rotate_matrix(w, h, angle):
    """Image.rotate(angle) (expand=0, centre = (w/2, h/2)): the inverse affine matrix Image.transform receives, or None
    when rotate() short-cuts (angle % 360 == 0 returns a copy)."""
    angle = angle % 360.0
    if angle == 0:
        return None
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def affine_fixed_coeffs(m):
    """The six 16.16 coefficients of Geometry.c affine_fixed (pixel centres folded into a2 / a5)."""
    return (_fix(m[0]), _fix(m[1]), _fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            _fix(m[3]), _fix(m[4]), _fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def rotate_nearest(img, angle):
    """Image.rotate(angle) with resample NEAREST, expand 0, fill 0.  img uint8 [h, w, 3]."""
    h, w = img.shape[:2]
    m = rotate_matrix(w, h, angle)
    if m is None:
        return img.copy()
    if (angle % 360.0) == 180:
        return img[::-1, ::-1].copy()
    if (angle % 360.0) in (90, 270) and w == h:
        return np.rot90(img, 1 if (angle % 360.0) == 90 else 3).copy()
    a0, a1, a2, a3, a4, a5 = affine_fixed_coeffs(m)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + ys * a1 + xs * a0) >> 16
    yin = (a5 + ys * a4 + xs * a3) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def scale_table(n_in, n_out):
    """Geometry.c ImagingScaleAffine index table of Image.resize(..., NEAREST): xo starts at a0 / 2 and is advanced by
    REPEATED double additions of a0 = n_in / n_out; COORD() truncates (negative -> -1)."""
    a0 = float(n_in) / float(n_out)
    tab = np.empty(n_out, np.int64)
    xo = 0.0 + a0 * 0.5
    for x in range(n_out):
        xin = -1 if xo < 0.0 else int(xo)
        tab[x] = xin if xin < n_in else -1
        xo += a0
    return tab


def resize_nearest(img, size):
    """Image.resize([W, H]) with resample NEAREST (Pillow 6.2.0's default)."""
    h, w = img.shape[:2]
    W, H = size
    xt, yt = scale_table(w, W), scale_table(h, H)
    out = np.zeros((H, W) + img.shape[2:], img.dtype)
    yy, xx = np.nonzero(yt >= 0)[0], np.nonzero(xt >= 0)[0]
    out[np.ix_(yy, xx)] = img[np.ix_(yt[yy], xt[xx])]
    return out


def crop_box(box):
    """Image.crop: every coordinate through int(round(.)) (Python 3 round: half to even)."""
    return tuple(int(round(v)) for v in box)


def crop(img, box):
    """Image.crop(box); the part of the box outside the image is zero."""
    x0, y0, x1, y1 = crop_box(box)
    h, w = img.shape[:2]
    out = np.zeros((max(y1 - y0, 0), max(x1 - x0, 0)) + img.shape[2:], img.dtype)
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = img[sy0:sy1, sx0:sx1]
    return out


def _clip8(v):
    """Filter.c clip8(float): <= 0 -> 0, >= 255 -> 255, else truncate."""
    return np.where(v <= 0.0, 0, np.where(v >= 255.0, 255, v.astype(np.int32))).astype(np.uint8)


def kernel_filter(img, fid):
    """Image.filter(ImageFilter.<built-in fid>) on RGB uint8: Filter.c ImagingFilter3x3 / 5x5.  float32 arithmetic: the
    kernel is divided by the scale in float32; per pixel ss = offset + 0.5, then one row of taps at a time is added,
    starting with the kernel's FIRST row on image row y+1 (Pillow applies the kernel bottom-up); the taps of a row are
    summed left to right.  The 1- (2-) pixel frame is copied from the input."""
    size, scale, kern = KERNELS[fid]
    k = (np.asarray(kern, np.float32) / np.float32(scale)).astype(np.float32).reshape(size, size)
    r = size // 2
    h, w = img.shape[:2]
    out = img.copy()
    if h <= 2 * r or w <= 2 * r:
        return out
    f = img.astype(np.float32)
    ss = np.full((h - 2 * r, w - 2 * r, img.shape[2]), np.float32(0.5), np.float32)
    for j in range(size):                   # kernel row j <-> image row y + r - j
        dy = r - j
        row = None
        for i in range(size):
            term = f[r + dy:h - r + dy, i:w - 2 * r + i] * k[j, i]
            row = term if row is None else (row + term).astype(np.float32)
        ss = (ss + row).astype(np.float32)
    out[r:h - r, r:w - r] = _clip8(ss)
    return out


def rgb_to_l(img):
    """Convert.c rgb2l: (R*19595 + G*38470 + B*7471 + 0x8000) >> 16."""
    v = img.astype(np.int64)
    return ((v[..., 0] * 19595 + v[..., 1] * 38470 + v[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(im1, im2, alpha):
    """Image.blend(im1, im2, alpha) (Blend.c): float32 alpha; in1 + alpha * (in2 - in1) in float32, truncated inside
    [0, 1], clipped + truncated outside."""
    a = np.float32(alpha)
    if a == 0.0:
        return im1.copy()
    if a == 1.0:
        return im2.copy()
    i1 = im1.astype(np.int32)
    d = (im2.astype(np.int32) - i1).astype(np.float32)
    t = (i1.astype(np.float32) + (a * d).astype(np.float32)).astype(np.float32)
    if 0.0 <= a <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return _clip8(t)


def enhance(img, fid, factor):
    """ImageEnhance.{Sharpness, Brightness, Color, Contrast}(img).enhance(factor)."""
    if fid == SHARPNESS:
        deg = kernel_filter(img, 2)
    elif fid == BRIGHTNESS:
        deg = np.zeros_like(img)
    elif fid == COLOR:
        deg = np.repeat(rgb_to_l(img)[..., None], 3, axis=2)
    elif fid == CONTRAST:
        lum = rgb_to_l(img)
        mean = int(float(lum.astype(np.int64).sum()) / float(lum.size) + 0.5)
        deg = np.full_like(img, mean)
    else:
        raise ValueError(fid)
    return blend(deg, img, factor)


def apply_filter(img, fid, r_val=0):
    """utils/data.py:8-35 for a drawn (r_id, r_val)."""
    if fid <= 5:
        return kernel_filter(img, fid)
    return enhance(img, fid, r_val * 0.1)


def to_model_range(img_u8):
    """``np.asarray(image) / 255.0`` (float64), cast to the dataset's tf.float32, then map_fn's ``* 2.0 - 1.0`` in fp32."""
    x = (img_u8.astype(np.float64) / 255.0).astype(np.float32)
    return (x * np.float32(2.0) - np.float32(1.0)).astype(np.float32)


def resized_size(w, h):
    """image_pair_dataloader.py:105-108 / 136-139: the short side becomes IMAGE_SIZE (int() of the float quotient)."""
    ratio = (h if w > h else w) / float(IMAGE_SIZE)
    return int(w / ratio), int(h / ratio), ratio


def pair_frame(img, plan):
    """One frame through image_pair_dataloader.py:95-159 for the drawn parameters ``plan`` (dict: randomness, angle, crop,
    flip, r_id, r_val)."""
    h, w = img.shape[:2]
    if plan["randomness"]:
        img = rotate_nearest(img, plan["angle"])
    W, H, _ = resized_size(w, h)
    img = resize_nearest(img, (W, H))
    if plan["randomness"]:
        c = plan["crop"]
        box = (c, 0, c + IMAGE_SIZE, IMAGE_SIZE) if w > h else (0, c, IMAGE_SIZE, c + IMAGE_SIZE)
        img = crop(img, box)
        if plan["flip"]:
            img = img[:, ::-1].copy()
        img = apply_filter(img, plan["r_id"], plan["r_val"])
    else:
        ox = W / 2.0            # both branches of the reference centre-crop along x (image_pair_dataloader.py:124-130,155-161)
        img = crop(img, (ox - IMAGE_SIZE // 2, 0, ox + IMAGE_SIZE // 2, IMAGE_SIZE))
    return img


def keypoint_frame(img, w0, h0):
    """One frame through keypoint_dataloader.py:66-72: resize by the FIRST frame's (w0, h0), centre crop of utils/data.py
    center_crop."""
    half = IMAGE_SIZE // 2
    W, H, _ = resized_size(w0, h0)
    if w0 > h0:
        box = (W / 2.0 - half, 0, W / 2.0 + half, IMAGE_SIZE)
    else:
        box = (0, H / 2.0 - half, IMAGE_SIZE, H / 2.0 + half)
    return crop(resize_nearest(img, (W, H)), box)
