"""The consumer of createGraph's data product: the graph network of the reference's ipt.py (ipt.py:28-84),
without DGL.

The reference builds a DGL graph with one node per triangle, an edge src -> dst for every non-zero
normalised transport weight w[dst, src] (edge feature = that weight), adds self loops (whose edge feature
is DGL's zero fill, so they carry nothing) and runs three message-passing layers
    reduced[dst] = sum over edges (src -> dst) of node_feats[src] * w[dst, src]
    node_feats   = act(Linear(cat(node_feats, reduced)))
between a tanh "lift" and a sigmoid output layer; node inputs are the eye row's pixel colours
(pixel[nT], the mean target colour seen through each triangle), labels the true albedos.

With nT ~ 30 nodes the message pass is a dense [nT, nT] @ [nT, hidden] product, so a scene is
(W, x): the thresholded, row-normalised weight matrix and the node inputs; batches are block-diagonal,
i.e. a batched matmul.  Plain torch: this is the learning pipeline around the hot path, not part of it.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

P_MIN = 1e-3   # ipt.py:26


def build_graph(w, pixel, light=None):
    """ipt.py:67-82 on the (w, pixel, light) arrays of generate_data: returns (W [nT, nT], x [nT, 3]).
    W[dst, src] is the weight of edge src -> dst; rows whose weights are all below P_MIN stay zero."""
    if torch.is_tensor(w):   # the device-resident product of TorchScene.graph_data: same steps, no host round trip
        w = torch.where(w < P_MIN, torch.zeros_like(w), w).to(torch.float64)
        s = w.sum(dim=-1, keepdim=True)
        w = w / torch.where(s != 0, s, torch.ones_like(s))
        return w[:-1].to(torch.float32), pixel[-1].to(torch.float32)
    w = np.array(w, np.float64, copy=True)
    w[w < P_MIN] = 0.0
    s = w.sum(axis=-1, keepdims=True)
    w = w / np.where(s != 0, s, np.ones_like(s))
    W, x = w[:-1], np.asarray(pixel)[-1]
    return torch.tensor(W, dtype=torch.float32), torch.tensor(x, dtype=torch.float32)


class MPL(nn.Module):
    """one message-passing layer (ipt.py:52-65)"""

    def __init__(self, in_feats, out_feats, activation=None):
        super().__init__()
        self.linear = nn.Linear(in_feats * 2, out_feats)
        self.activation = activation

    def forward(self, W, h):
        reduced = torch.matmul(W, h)
        h = self.linear(torch.cat((h, reduced), dim=-1))
        return self.activation(h) if self.activation is not None else h


class GCN(nn.Module):
    """ipt.py:28-50: Linear+tanh lift, three MPL(relu), Linear+sigmoid -> per-triangle albedo in (0, 1)"""

    def __init__(self, in_feats=3, out_feats=3, hidden=100):
        super().__init__()
        self.lift = nn.Linear(in_feats, hidden)
        self.network = nn.ModuleList([MPL(hidden, hidden, F.relu) for _ in range(3)])
        self.out = nn.Linear(hidden, out_feats)

    def forward(self, W, x):
        """W: [nT, nT] or [B, nT, nT]; x: [nT, in] or [B, nT, in]"""
        h = self.lift(x).tanh()
        for layer in self.network:
            h = layer(W, h)
        return self.out(h).sigmoid()

    @staticmethod
    def loss(preds, labels):
        return (preds - labels).abs().mean()


def fit(data, epochs=1000, lr=1e-4, device=None, seed=0, log_every=0):
    """ipt.py:84-125: Adam on the mean absolute albedo error over `data`, a list of generate_data tuples
    (w, pixel, light, labels) from scenes with the same triangle count.  Returns (model, final loss)."""
    device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    torch.manual_seed(seed)
    graphs = [build_graph(w, pixel, light) for (w, pixel, light, _) in data]
    W = torch.stack([g[0] for g in graphs]).to(device)
    x = torch.stack([g[1] for g in graphs]).to(device)
    y = torch.stack([torch.as_tensor(np.asarray(lbl), dtype=torch.float32) for (_, _, _, lbl) in data]).to(device)
    model = GCN(x.shape[-1], y.shape[-1]).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    last = None
    for i in range(epochs):
        loss = model.loss(model(W, x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        last = float(loss.detach())
        if log_every and (i + 1) % log_every == 0:
            print(i + 1, last)
    return model, last
