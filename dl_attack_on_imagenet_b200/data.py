"""Synthetic data / classifier helpers used by bench, smoke and tests (the reference's ImageNet I/O --
DS_ImageNet.py, imagenet_loading.py -- is out of scope; only the `.indexed` dataset protocol is kept)."""
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class IndexedTensorDataset(torch.utils.data.Dataset):
    """Tensor-backed dataset speaking the protocol of imagenet_loading.Subset_I (imagenet_loading.py:8-18):
    `.indexed = True` makes __getitem__ return (item, x, y) instead of (x, y).  Exposes `.images` / `.labels`
    so that ADIL can keep the whole set resident in HBM and gather rows inside the synthesis kernel."""

    def __init__(self, images, labels, indexed=False):
        self.images = images
        self.labels = labels
        self.indexed = indexed

    def __len__(self):
        return len(self.images)

    def __getitem__(self, item):
        if self.indexed:
            return item, self.images[item], self.labels[item]
        return self.images[item], self.labels[item]


class Normalize(torch.nn.Module):
    """(input - mean) / std per channel -- demo_dL_attack.py:16-25.  ADIL recognises this module (any leading
    module with 3-element `mean` / `std` buffers) and fuses it into the synthesis / backward kernels."""

    def __init__(self, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        super().__init__()
        self.register_buffer('mean', torch.tensor(mean, dtype=torch.float32))
        self.register_buffer('std', torch.tensor(std, dtype=torch.float32))

    def forward(self, input):
        mean = self.mean.reshape(1, -1, 1, 1)
        std = self.std.reshape(1, -1, 1, 1)
        return (input - mean) / std


_ZOO = {'resnet': 'resnet18', 'resnet18': 'resnet18', 'resnet50': 'resnet50', 'densenet': 'densenet121',
        'densenet121': 'densenet121', 'googlenet': 'googlenet', 'inception': 'inception_v3',
        'inception_v3': 'inception_v3', 'mobilenet': 'mobilenet_v2', 'mobilenet_v2': 'mobilenet_v2',
        'vgg': 'vgg11', 'vgg11': 'vgg11', 'vgg16': 'vgg16'}


def build_classifier(name, seed=0, device='cpu'):
    """Random-init torchvision classifier wrapped as Sequential(Normalize, net) like demo_dL_attack.py:41-59
    (weights=None: there is no network for checkpoints)."""
    import torchvision.models as models
    arch = _ZOO[name.lower()]
    torch.manual_seed(seed)
    kwargs = {'init_weights': True} if arch in ('googlenet', 'inception_v3') else {}
    net = getattr(models, arch)(weights=None, **kwargs)
    return torch.nn.Sequential(Normalize(), net).eval().to(device)


def synthetic_images(n, seed, size=224, channels=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, channels, size, size, generator=g)
    y = torch.randint(0, 1000, (n,), generator=g)
    return x, y


class HostBatchPrefetcher(object):
    """Host -> device staging of minibatches for `ADIL.fit_batch` when the image set lives in host memory (the
    reference's DataLoader path, adil.py:130,168): the rows of batch i+1 are gathered into pinned memory and copied
    on a side stream while the GPU computes batch i.  Every batch is still copied from pinned host memory; the copy
    just no longer serialises with the kernels.

        pf = HostBatchPrefetcher(x_host, device)
        pf.submit(idx0)
        for i in range(steps):
            x_dev = pf.get()                      # current stream waits for the copy of batch i
            loss, fooled = attack.fit_batch(idx[i], x_dev)
            pf.release()                          # batch i consumed (its device buffer may be refilled later)
            if i + 1 < steps: pf.submit(idx[i + 1])
    """

    def __init__(self, x_host, device, depth=2):
        self.x_host = x_host
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._pinned, self._dev = [None] * depth, [None] * depth
        self._copied = [torch.cuda.Event() for _ in range(depth)]     # H2D of slot k finished
        self._consumed = [None] * depth                               # compute that read slot k was enqueued
        self._head = 0   # next slot to fill
        self._tail = 0   # next slot to hand out
        self._inflight = 0

    def submit(self, index):
        """Gather x_host[index] into pinned memory and start its copy to the device (returns at once)."""
        if self._inflight >= self.depth:
            raise RuntimeError("HostBatchPrefetcher: all %d slots are in flight; call get()/release() first" % self.depth)
        k = self._head
        index = torch.as_tensor(index, dtype=torch.long)
        shape = (index.numel(),) + tuple(self.x_host.shape[1:])
        if self._pinned[k] is None or tuple(self._pinned[k].shape) != shape:
            self._pinned[k] = torch.empty(shape, dtype=self.x_host.dtype, pin_memory=True)
            self._dev[k] = torch.empty(shape, dtype=self.x_host.dtype, device=self.device)
        else:
            self._copied[k].synchronize()          # the previous copy out of this pinned buffer has finished
        torch.index_select(self.x_host, 0, index, out=self._pinned[k])
        with torch.cuda.stream(self.copy_stream):
            if self._consumed[k] is not None:
                self.copy_stream.wait_event(self._consumed[k])    # the kernels that read this device buffer are done
            self._dev[k].copy_(self._pinned[k], non_blocking=True)
            self._copied[k].record(self.copy_stream)
        self._head = (k + 1) % self.depth
        self._inflight += 1

    def get(self):
        """Device tensor of the oldest submitted batch; the current stream waits for its copy."""
        if self._inflight == 0:
            raise RuntimeError("HostBatchPrefetcher: nothing submitted")
        k = self._tail
        torch.cuda.current_stream(self.device).wait_event(self._copied[k])
        return self._dev[k]

    def release(self):
        """Mark the batch returned by the last get() as consumed by the work enqueued so far."""
        k = self._tail
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._consumed[k] = ev
        self._tail = (k + 1) % self.depth
        self._inflight -= 1

    @property
    def bytes_per_batch(self):
        b = self._pinned[(self._head - 1) % self.depth]
        return 0 if b is None else b.numel() * b.element_size()
