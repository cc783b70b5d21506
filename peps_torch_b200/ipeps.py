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


# ----------------------------------------------------------------------------------------------
# state I/O: the reference's JSON files (ipeps/ipeps.py:339-441 read_ipeps, :501-535 write_ipeps; tensor formats of
# ipeps/tensor_io.py:46-100 and :186-247) -- what `--instate` / `--out_prefix` of the example scripts read and write.
# Host-side; the tensors are parsed with numpy and land on the device in one copy each.
# ----------------------------------------------------------------------------------------------
def _tensor_from_json(t):
    import numpy as np
    dtype_str = t.get("dtype", "float64").lower()
    if dtype_str not in ("float64", "complex128"):
        raise ValueError("Invalid dtype " + dtype_str)
    if t.get("format") == "1D":
        data = np.asarray(t["data"], dtype=np.complex128 if "complex" in dtype_str else np.float64)
        return data.reshape(t["dims"])
    dims = t["dims"] if "dims" in t else [t["physDim"]] + [t["auxDim"]] * 4
    X = np.zeros(dims, dtype=dtype_str)
    nd = len(dims)
    if t["entries"]:
        rows = np.array([e.split() for e in t["entries"]])
        idx = tuple(rows[:, i].astype(np.int64) for i in range(nd))
        vals = rows[:, nd].astype(np.float64)
        if dtype_str == "complex128":
            vals = vals + 1j * rows[:, nd + 1].astype(np.float64)
        X[idx] = vals
    return X


def _tensor_to_json(t, fmt):
    import numpy as np
    a = t.detach().cpu().numpy()
    dtype_str = str(a.dtype)
    out = {"dtype": dtype_str, "dims": list(a.shape)}
    if fmt == "1D":
        out["format"] = "1D"
        out["data"] = [repr(x.item()) if not np.iscomplexobj(a) else str(x.item()) for x in a.reshape(-1)]
        return out
    entries = []
    for ei in np.ndindex(*a.shape):
        v = a[ei]
        head = " ".join(str(i) for i in ei)
        entries.append(f"{head} {v.real!r} {v.imag!r}" if np.iscomplexobj(a) else f"{head} {float(v)!r}")
    out["numEntries"] = len(entries)
    out["entries"] = entries
    return out


def read_ipeps(jsonfile, vertexToSite=None, aux_seq=(0, 1, 2, 3), dtype=None, device='cpu'):
    """An IPEPS from the reference's JSON format.  `aux_seq`: order of the auxiliary indices in the file relative to
    [up, left, down, right] (overridden by the file's own "aux_ind_seq"); `dtype` complex128 promotes real tensors."""
    import json
    import torch
    asq = [x + 1 for x in aux_seq]
    with open(jsonfile) as f:
        raw = json.load(f)
    if "aux_ind_seq" in raw:
        asq = [x + 1 for x in raw["aux_ind_seq"]]
    by_id = {s["siteId"]: s for s in raw["sites"]}
    sites = OrderedDict()
    for ts in raw["map"]:
        if ts["siteId"] not in by_id:
            raise Exception("Tensor with siteId: " + ts["siteId"] + " NOT FOUND in \"sites\"")
        X = torch.from_numpy(_tensor_from_json(by_id[ts["siteId"]])).permute((0, *asq)).contiguous()
        if dtype is not None and dtype.is_complex and not X.is_complex():
            X = X + 0.j
        sites[(ts["x"], ts["y"])] = X.to(device)
    lX = raw["sizeM"] if "sizeM" in raw else raw["lX"]
    lY = raw["sizeN"] if "sizeN" in raw else raw["lY"]
    if vertexToSite is None:
        if "pattern" in raw:
            # pattern[y][x] = siteId of the tensor at (x, y) (ipeps/ipeps.py:186-207)
            pattern = raw["pattern"]
            id2coord = {ts["siteId"]: (ts["x"], ts["y"]) for ts in raw["map"]}
            pY, pX = len(pattern), len(pattern[0])

            def vertexToSite(coord):
                return id2coord[pattern[coord[1] % pY][coord[0] % pX]]
        else:
            def vertexToSite(coord):
                return ((coord[0] + abs(coord[0]) * lX) % lX, (coord[1] + abs(coord[1]) * lY) % lY)
    return IPEPS(sites, vertexToSite, lX=lX, lY=lY)


def write_ipeps(state, outputfile, aux_seq=(0, 1, 2, 3), normalize=False, tensor_io_format="legacy"):
    """The reference's JSON format (legacy entry lists by default, as config.py's tensor_io_format)."""
    import json
    asq = [x + 1 for x in aux_seq]
    out = {"lX": state.lX, "lY": state.lY, "sites": []}
    ids, smap = [], []
    for nid, (coord, site) in enumerate(state.sites.items()):
        if normalize:
            site = site / site.abs().max()
        ids.append(f"A{nid}")
        smap.append({"siteId": ids[-1], "x": coord[0], "y": coord[1]})
        jt = _tensor_to_json(site.permute((0, *asq)), tensor_io_format)
        jt["siteId"] = ids[-1]
        out["sites"].append(jt)
    out["siteIds"] = ids
    out["map"] = smap
    if tuple(aux_seq) != (0, 1, 2, 3):
        out["aux_ind_seq"] = list(aux_seq)
    c2id = {(r["x"], r["y"]): r["siteId"] for r in smap}
    out["pattern"] = [[c2id[state.vertexToSite((x, y))] for x in range(state.lX)] for y in range(state.lY)]
    with open(outputfile, 'w') as f:
        json.dump(out, f, indent=4, separators=(',', ': '))
