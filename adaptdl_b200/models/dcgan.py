"""DCGAN generator / discriminator (Radford et al. 2015) -- the reference's
``examples/dcgan`` workload: two networks, two optimizers, two
``AdaptiveDataParallel`` instances whose gradient statistics are summed."""

import torch.nn as nn

__all__ = ["Generator", "Discriminator", "weights_init"]


def weights_init(m):
    name = m.__class__.__name__
    if name.find("Conv") != -1:
        nn.init.normal_(m.weight, 0.0, 0.02)
    elif name.find("BatchNorm") != -1:
        nn.init.normal_(m.weight, 1.0, 0.02)
        nn.init.zeros_(m.bias)


class Generator(nn.Module):
    def __init__(self, nz=100, ngf=64, nc=3):
        super().__init__()

        def up(cin, cout, k, s, p):
            return [nn.ConvTranspose2d(cin, cout, k, s, p, bias=False),
                    nn.BatchNorm2d(cout), nn.ReLU(True)]
        self.main = nn.Sequential(
            *up(nz, ngf * 8, 4, 1, 0), *up(ngf * 8, ngf * 4, 4, 2, 1),
            *up(ngf * 4, ngf * 2, 4, 2, 1), *up(ngf * 2, ngf, 4, 2, 1),
            nn.ConvTranspose2d(ngf, nc, 4, 2, 1, bias=False), nn.Tanh())
        self.apply(weights_init)

    def forward(self, z):
        return self.main(z)


class Discriminator(nn.Module):
    def __init__(self, ndf=64, nc=3):
        super().__init__()

        def down(cin, cout, bn=True):
            layers = [nn.Conv2d(cin, cout, 4, 2, 1, bias=False)]
            if bn:
                layers.append(nn.BatchNorm2d(cout))
            return layers + [nn.LeakyReLU(0.2, inplace=True)]
        self.main = nn.Sequential(
            *down(nc, ndf, bn=False), *down(ndf, ndf * 2),
            *down(ndf * 2, ndf * 4), *down(ndf * 4, ndf * 8),
            nn.Conv2d(ndf * 8, 1, 4, 1, 0, bias=False))
        self.apply(weights_init)

    def forward(self, x):
        return self.main(x).view(-1)
