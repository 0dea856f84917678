"""Host-side image transforms of the detector scripts (reference train.py:118-164: `transforms.Compose([...])` from
mxnet.gluon.data.vision) for frames decoded on the CPU.

The geometric part stays on the host, on uint8 HWC images (OpenCV, the backend MXNet's image ops use as well); `ToTensor` and
`Normalize` are NOT done here: batches go to the GPU as uint8 NHWC — a quarter of the fp32 bytes over PCIe — and the backbone's
input kernel applies /255 and (x - mean) / std while it builds the space-to-depth image (`TN_FRAMES_U8_NHWC`).

  test  : Resize(S + 32) -> CenterCrop(S)                       (an int size resizes to a SQUARE, aspect not kept: SURVEY.md A.1)
  train : RandomResizedCrop(S) -> RandomFlipLeftRight -> RandomColorJitter(0.4, 0.4, 0.4) -> RandomLighting(0.1)
          (only for window == 1: clips, feature dumps and temporal pooling reuse the test transform, train.py:159-164)
"""
import numpy as np
import torch


def _cv2():
    import cv2
    return cv2


def _as_numpy(img):
    a = img.numpy() if isinstance(img, torch.Tensor) else np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected a uint8 HWC RGB image, got %s %s" % (a.dtype, a.shape))
    return a


def resize(img, size):
    """mx Resize(size) with an int: (size, size), bilinear."""
    return _cv2().resize(_as_numpy(img), (int(size), int(size)), interpolation=_cv2().INTER_LINEAR)


def center_crop(img, size):
    a = _as_numpy(img)
    h, w = a.shape[:2]
    if h < size or w < size:
        raise ValueError("image %dx%d is smaller than the crop %d" % (h, w, size))
    y0, x0 = int((h - size) / 2), int((w - size) / 2)
    return a[y0:y0 + size, x0:x0 + size]


_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def to_tensor_normalize(a):
    """ToTensor + Normalize on the host (train.py:138-139,145-146): uint8 HWC -> float32 CHW, (x/255 - mean) / std.  Used when
    the frames feed the fp32 training graph of a trainable backbone; inference batches stay uint8 and are normalised on the GPU."""
    x = (a.astype(np.float32) / 255.0 - _MEAN) / _STD
    return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1)))


class TestTransform(object):
    """Resize(S + 32) -> CenterCrop(S); returns a uint8 (S,S,3) torch tensor, or float32 (3,S,S) with `to_tensor`."""

    def __init__(self, data_shape, to_tensor=False):
        self.S, self.to_tensor = int(data_shape), to_tensor

    def __call__(self, img):
        a = np.ascontiguousarray(center_crop(resize(img, self.S + 32), self.S))
        return to_tensor_normalize(a) if self.to_tensor else torch.from_numpy(a)


class TrainTransform(object):
    """Training-time augmentation for single frames; `seed` makes it reproducible."""

    # ImageNet PCA lighting (AlexNet), the constants behind RandomLighting
    _EIGVAL = np.array([55.46, 4.794, 1.148], dtype=np.float32)
    _EIGVEC = np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]], dtype=np.float32)

    def __init__(self, data_shape, jitter=0.4, lighting=0.1, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), seed=None,
                 to_tensor=False):
        self.S, self.jitter, self.lighting, self.scale, self.ratio = int(data_shape), jitter, lighting, scale, ratio
        self.to_tensor = to_tensor
        self.rng = np.random.RandomState(seed)

    def _random_resized_crop(self, a):
        h, w = a.shape[:2]
        area = h * w
        for _ in range(10):
            target = self.rng.uniform(*self.scale) * area
            log_r = self.rng.uniform(np.log(self.ratio[0]), np.log(self.ratio[1]))
            cw, ch = int(round(np.sqrt(target * np.exp(log_r)))), int(round(np.sqrt(target / np.exp(log_r))))
            if 0 < cw <= w and 0 < ch <= h:
                x0, y0 = self.rng.randint(0, w - cw + 1), self.rng.randint(0, h - ch + 1)
                return _cv2().resize(a[y0:y0 + ch, x0:x0 + cw], (self.S, self.S), interpolation=_cv2().INTER_LINEAR)
        return center_crop(resize(a, self.S + 32), self.S)

    def __call__(self, img):
        a = self._random_resized_crop(_as_numpy(img))
        if self.rng.rand() < 0.5:
            a = a[:, ::-1]
        x = a.astype(np.float32)
        if self.jitter > 0:
            gray_w = np.array([0.299, 0.587, 0.114], dtype=np.float32)
            ops = [0, 1, 2]
            self.rng.shuffle(ops)
            for op in ops:
                alpha = 1.0 + self.rng.uniform(-self.jitter, self.jitter)
                if op == 0:  # brightness
                    x = x * alpha
                elif op == 1:  # contrast: blend with the mean luminance
                    x = x * alpha + (1.0 - alpha) * float((x * gray_w).sum(axis=2).mean())
                else:  # saturation: blend with the per-pixel luminance
                    x = x * alpha + (1.0 - alpha) * (x * gray_w).sum(axis=2, keepdims=True)
        if self.lighting > 0:
            a3 = self.rng.normal(0, self.lighting, size=3).astype(np.float32)
            x = x + (self._EIGVEC * a3 * self._EIGVAL).sum(axis=1)
        a = np.ascontiguousarray(np.clip(x, 0, 255).astype(np.uint8))
        return to_tensor_normalize(a) if self.to_tensor else torch.from_numpy(a)


def build_transforms(data_shape, window=1, save_feats=False, to_tensor=False):
    """(train, test) transforms as the reference assembles them (train.py:118-164).  `to_tensor`: normalised fp32 CHW output
    (for the fp32 training graph of a trainable backbone) instead of uint8 HWC."""
    test = TestTransform(data_shape, to_tensor=to_tensor)
    train = test if (window > 1 or save_feats) else TrainTransform(data_shape, to_tensor=to_tensor)
    return train, test
