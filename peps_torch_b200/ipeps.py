"""Minimal state containers with the attributes the CTM move reads from the reference's
IPEPS / IPEPS_C4V (ipeps/ipeps.py:89-250, ipeps/ipeps_c4v.py:6-33).  The engine is duck-typed:
the reference's own objects work unchanged; these exist so that tests and bench run on a box
where the reference is absent."""
from collections import OrderedDict


class IPEPS:
    def __init__(self, sites, vertexToSite=None, lX=None, lY=None):
        self.sites = OrderedDict(sites)
        self.vertexToSite = vertexToSite if vertexToSite is not None else (lambda coord: (0, 0))
        xs = sorted({c[0] for c in self.sites})
        ys = sorted({c[1] for c in self.sites})
        self.lX = lX if lX is not None else len(xs)
        self.lY = lY if lY is not None else len(ys)
        t = next(iter(self.sites.values()))
        self.dtype, self.device = t.dtype, t.device

    def site(self, coord):
        return self.sites[self.vertexToSite(coord)]

    def get_parameters(self):
        """The on-site tensors (ipeps/ipeps.py:249-256)."""
        return self.sites.values()

    def get_aux_bond_dims(self):
        """All auxiliary bond dimensions of all sites, site by site in (u, l, d, r) order (ipeps/ipeps.py:306-307)."""
        return [d for t in self.sites.values() for d in t.shape[1:]]

    def normalize_(self):
        """Every on-site tensor divided by its largest magnitude (ipeps/ipeps.py:331-333)."""
        for c in self.sites:
            self.sites[c] = self.sites[c] / self.sites[c].abs().max()

    def __str__(self):
        lines = [f"lX x lY: {self.lX} x {self.lY}"]
        lines += [f"a{i} {c}: {tuple(t.shape)}" for i, (c, t) in enumerate(self.sites.items())]
        return "\n".join(lines)

    def to(self, device):
        return IPEPS(OrderedDict((c, t.to(device)) for c, t in self.sites.items()), self.vertexToSite, self.lX, self.lY)


class IPEPS_C4V(IPEPS):
    def __init__(self, site):
        super().__init__(OrderedDict({(0, 0): site}), lambda coord: (0, 0), 1, 1)

    def site(self, coord=None):
        return self.sites[(0, 0)]
