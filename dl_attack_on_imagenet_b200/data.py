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
