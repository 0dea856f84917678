"""Device-side accumulation of the detector's metrics (reference metrics/vision.py:27-58 PRF1 + mx.metric.Accuracy / top-k,
updated from a device->host copy of every batch's logits at train.py:427-431, 503-527 and evaluate.py:274-303).

The counters (confusion matrix, top-1 / top-k hits) live on the GPU and one kernel (tn_metrics_update) folds a batch in, so the
evaluation loop has no per-batch synchronisation; `finish()` reads them once, sums them over the ranks when torch.distributed is
initialised, and returns the same host metric objects the scripts print from."""
import numpy as np
import torch

from .._lib import check, dptr, lib, stream_ptr
from .vision import PRF1, Accuracy


class DeviceMetrics(object):
    def __init__(self, label_names, top_k=5, device=None):
        self.label_names = list(label_names)
        self.top_k = int(top_k)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n = len(self.label_names)
        # uint64 counters, stored in int64 tensors (torch has no uint64 arithmetic; the values are counts)
        self._conf = torch.zeros(n, n, dtype=torch.int64, device=self.device)
        self._hits = torch.zeros(3, dtype=torch.int64, device=self.device)

    def reset(self):
        self._conf.zero_()
        self._hits.zero_()

    def update(self, labels, logits):
        """labels (N,) any integer dtype, logits (N,C) fp32 -- both already on the device; nothing is copied to the host."""
        if not logits.is_cuda:
            raise ValueError("DeviceMetrics.update takes device tensors (host-side metrics: tennis_b200.metrics.vision)")
        n, c = logits.shape
        assert c == len(self.label_names)
        lab = labels.to(device=logits.device, dtype=torch.int32).contiguous()
        lg = logits.float().contiguous()
        check(lib().tn_metrics_update(dptr(lg), dptr(lab), n, c, self.top_k, dptr(self._conf), dptr(self._hits), stream_ptr()))

    def finish(self):
        """-> [Accuracy(top-1), Accuracy('top%d' % k), PRF1] filled from the device counters (summed over ranks)."""
        import torch.distributed as dist
        conf, hits = self._conf, self._hits
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            conf, hits = conf.clone(), hits.clone()
            dist.all_reduce(conf)
            dist.all_reduce(hits)
        conf = conf.cpu().numpy().astype(np.float64)
        hits = hits.cpu().numpy()
        acc, topk, prf = Accuracy(), Accuracy('top%d' % self.top_k, top_k=self.top_k), PRF1(label_names=self.label_names)
        acc.hit, acc.n = int(hits[0]), int(hits[2])
        topk.hit, topk.n = int(hits[1]), int(hits[2])
        prf.mat = conf
        prf.scores = np.stack([np.diag(conf), conf.sum(axis=1), conf.sum(axis=0)])  # matches, positives (labels), predictions
        return [acc, topk, prf]
