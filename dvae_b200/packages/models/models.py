"""VAE containers with the reference's class names, constructor signatures and state_dict layout, whose forward
passes run on libdvae_b200's CUDA kernels.

Drop-in for ``packages/models/models.py`` as far as the enhancement / reconstruction scripts use it:
``VariationalAutoencoder`` (125-182), ``DeepGenerativeModel`` (185-218), ``DeepGenerativeModel_v3`` (245-297),
``DeepGenerativeModel_v5`` (390-444, used through ``.enc_dec_clf``), built from ``Encoder`` (91-105), ``Decoder``
(108-122), ``GaussianSample`` (24-38) and ``Classifier`` (41-63; parameters only, so checkpoints load).

The modules own ordinary ``nn.Linear`` parameters (same names, same xavier-normal / zero-bias initialisation), so
``load_state_dict(torch.load(...))`` of a reference checkpoint works unchanged.  ``forward`` needs CUDA tensors: the
layers are packed once into transposed FP32 buffers and evaluated by ``dvae_mlp_fwd``; there is no CPU path.
Training-time pieces of the reference file (KL terms, flows, ``_v2`` / ``_v4`` variants, the classifiers' forward) are
out of scope of this library.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import init

from ... import _lib
from ...engine import PackedMlp, mlp_forward, _p, _stream


def _xavier(module: nn.Module):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            init.xavier_normal_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()


class _PackedModule(nn.Module):
    """Caches the device-side packed copy of its linear layers; re-packs when a parameter changes or moves."""

    def _packed(self, name, layer_lists):
        key = tuple((p.data_ptr(), p._version, str(p.device)) for layers in layer_lists for l in layers for p in (l.weight, l.bias))
        cache = self.__dict__.setdefault("_pack_cache", {})
        hit = cache.get(name)
        if hit is None or hit[0] != key:
            dev = layer_lists[0][0].weight.device
            if dev.type != "cuda":
                raise _lib.DvaeError("dvae_b200 modules run on CUDA only: call .to(device) first (no CPU fallback)")
            packs = [PackedMlp([(l.weight.data, l.bias.data) for l in layers], dev) for layers in layer_lists]
            cache[name] = hit = (key, packs)
        return hit[1]

    def __getstate__(self):                      # keep pickling (spawn pools) free of device handles
        d = self.__dict__.copy()
        d.pop("_pack_cache", None)
        return d


def _rows(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise _lib.DvaeError("dvae_b200 modules take CUDA tensors (no CPU fallback)")
    x = x.detach().to(torch.float32)
    return x if x.stride(-1) == 1 else x.contiguous()


class GaussianSample(_PackedModule):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.mu = nn.Linear(in_features, out_features)
        self.log_var = nn.Linear(in_features, out_features)


class Encoder(_PackedModule):
    """tanh MLP + Gaussian head: ``forward(x) -> (z, mu, log_var)`` (models.py:91-105)."""

    def __init__(self, dims, sample_layer=GaussianSample):
        super().__init__()
        x_dim, h_dim, z_dim = dims
        neurons = [x_dim, *h_dim]
        self.hidden = nn.ModuleList([nn.Linear(neurons[i - 1], neurons[i]) for i in range(1, len(neurons))])
        self.sample = sample_layer(h_dim[-1], z_dim)

    def forward(self, x):
        x = _rows(x)
        shape = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        trunk = list(self.hidden)
        mu_p, lv_p = self._packed("enc", [trunk + [self.sample.mu], trunk + [self.sample.log_var]])
        mu = mlp_forward(mu_p, x2, _lib.ACT_NONE)
        log_var = mlp_forward(lv_p, x2, _lib.ACT_NONE)
        # the reference draws the reparametrisation noise on the CPU generator and moves it (models.py:10-13)
        eps = torch.randn(mu.size(), requires_grad=False).to(mu.device, non_blocking=True)
        z = torch.empty_like(mu)
        _lib.call("dvae_reparam", _p(mu), _p(log_var), _p(eps), _p(z), mu.numel(), _stream())
        L = mu.shape[-1]
        return z.reshape(*shape, L), mu.reshape(*shape, L), log_var.reshape(*shape, L)


class Decoder(_PackedModule):
    """tanh MLP with exponential output: ``forward(z) -> exp(...)`` (models.py:108-122)."""

    def __init__(self, dims):
        super().__init__()
        z_dim, h_dim, x_dim = dims
        neurons = [z_dim, *h_dim]
        self.hidden = nn.ModuleList([nn.Linear(neurons[i - 1], neurons[i]) for i in range(1, len(neurons))])
        self.reconstruction = nn.Linear(h_dim[-1], x_dim)

    def forward(self, x):
        x = _rows(x)
        shape = x.shape[:-1]
        (pack,) = self._packed("dec", [list(self.hidden) + [self.reconstruction]])
        out = mlp_forward(pack, x.reshape(-1, x.shape[-1]), _lib.ACT_EXP)
        return out.reshape(*shape, out.shape[-1])


class Classifier(nn.Module):
    """Parameter container only (models.py:41-63): keeps ``classifier.*`` / ``auxiliary.*`` checkpoint keys loadable."""

    def __init__(self, dims, batch_norm=False):
        super().__init__()
        x_dim, h_dim, y_dim = dims
        neurons = [x_dim, *h_dim]
        layers = []
        for i in range(1, len(neurons)):
            layers.append(nn.Linear(neurons[i - 1], neurons[i]))
            if batch_norm:
                layers.append(nn.BatchNorm1d(neurons[i]))
        self.hidden = nn.ModuleList(layers)
        self.output_layer = nn.Linear(h_dim[-1], y_dim)

    def forward(self, x):
        raise NotImplementedError("classifier inference is outside the enhancement path served by dvae_b200")


class VariationalAutoencoder(nn.Module):
    """M1 audio-only VAE: ``VariationalAutoencoder([x_dim, z_dim, h_dim])`` (models.py:125-182)."""

    def __init__(self, dims):
        super().__init__()
        x_dim, z_dim, h_dim = dims
        self.z_dim = z_dim
        self.flow = None
        self.encoder = Encoder([x_dim, h_dim, z_dim])
        self.decoder = Decoder([z_dim, list(reversed(h_dim)), x_dim])
        self.kl_divergence = 0
        _xavier(self)

    def forward(self, x, y=None):
        z, z_mu, z_log_var = self.encoder(x)
        self.kl_divergence = -0.5 * torch.sum(z_log_var - z_mu.pow(2) - z_log_var.exp(), dim=-1)   # _kld_v2
        return self.decoder(z), z_mu, z_log_var

    def sample(self, z):
        return self.decoder(z)


class DeepGenerativeModel(VariationalAutoencoder):
    """M2: encoder over ``[x; y]``, decoder over ``[z; y]`` (models.py:185-218)."""

    def __init__(self, dims, classifier):
        x_dim, self.y_dim, z_dim, h_dim = dims
        super().__init__([x_dim, z_dim, h_dim])
        self.encoder = Encoder([x_dim + self.y_dim, h_dim, z_dim])
        self.decoder = Decoder([z_dim + self.y_dim, list(reversed(h_dim)), x_dim])
        self.classifier = classifier
        _xavier(self)

    def forward(self, x, y):
        z, z_mu, z_log_var = self.encoder(torch.cat([x, y], dim=1))
        return self.decoder(torch.cat([z, y], dim=1)), z_mu, z_log_var

    def sample(self, z, y):
        return self.decoder(torch.cat([z, y.float()], dim=1))


class DeepGenerativeModel_v3(nn.Module):
    """M2-info core: encoder over ``x`` only, decoder over ``[z; y]``, plus a classifier head (models.py:245-297)."""

    def __init__(self, dims):
        x_dim, self.y_dim, z_dim, h_dim = dims
        self.z_dim = z_dim
        self.flow = None
        super().__init__()
        self.encoder = Encoder([x_dim, h_dim, z_dim])
        self.decoder = Decoder([z_dim + self.y_dim, list(reversed(h_dim)), x_dim])
        self.classifier = Classifier([x_dim, h_dim, self.y_dim])
        _xavier(self)

    def forward(self, x, y):
        z, z_mu, z_log_var = self.encoder(x)
        return self.decoder(torch.cat([z, y], dim=1)), z_mu, z_log_var

    def sample(self, z, y):
        return self.decoder(torch.cat([z, y.float()], dim=1))


class DeepGenerativeModel_v5(nn.Module):
    """M2-info with the adversarial auxiliary classifier (models.py:390-444); evaluation uses ``.enc_dec_clf``."""

    def __init__(self, dims):
        x_dim, self.y_dim, z_dim, h_dim = dims
        super().__init__()
        self.enc_dec_clf = DeepGenerativeModel_v3([x_dim, self.y_dim, z_dim, h_dim])
        self.auxiliary = Classifier([z_dim, h_dim, self.y_dim])
        _xavier(self)

    def forward(self, x, y):
        z, z_mu, z_log_var = self.enc_dec_clf.encoder(x)
        return self.enc_dec_clf.decoder(torch.cat([z, y], dim=1)), z, z_mu, z_log_var

    def sample(self, z, y):
        return self.enc_dec_clf.decoder(torch.cat([z, y.float()], dim=1))
