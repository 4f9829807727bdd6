"""Width tables of the LINNA emulator networks, as plain data.

The reference builds its fully-connected emulators in ``linna/nn.py``:
``ChtoModelv2`` (:59-133), ``ChtoModelv2_linear`` (:136-198) and
``ChtoModelsimple`` (:300-374), all sharing ``ResBlock_batchnorm`` (:11-56).
This module restates only their *shapes* and state-dict key names so that the
packer (``engine.py``), the synthetic-problem factory and the C oracle agree on
one op list.  No arithmetic lives here.

An op list is a list of dicts:

``{"kind": "linear", "name": "layer1", "in": K, "out": N, "act": "relu"|"none"}``
``{"kind": "res", "name": "layer2", "in": K, "mid": C, "out": N, "alpha": 0.1}``

For ``res``:  h = relu(W1 x + b1);  y = relu(alpha*(W2 h + b2) + Ws x)
(reference ``linna/nn.py:53-54``); ``Ws`` is the identity when in == out
(``linna/nn.py:28-31``).
"""

MODEL_KINDS = ("ChtoModelv2", "ChtoModelv2_linear", "ChtoModelsimple")


def hidden_size(n_out):
    """First hidden width, ``linna/nn.py:73-76``."""
    h = max(32, int(n_out * 32))
    if n_out > 30:
        h = 1000
    return h


def chto_ops(kind, n_in, n_out):
    """Op list for one of the three reference model classes."""
    if kind not in MODEL_KINDS:
        raise ValueError("unknown model kind %r" % (kind,))
    channel = 4 if kind == "ChtoModelsimple" else 16
    h = hidden_size(n_out)
    ops = [dict(kind="linear", name="layer1", **{"in": n_in, "out": h, "act": "relu"})]
    for name, mult in (("layer2", 1), ("layer3", 2), ("layer4", 4)):
        ops.append(dict(kind="res", name=name, **{"in": h, "mid": int(channel * mult),
                                                   "out": h // 2, "alpha": 0.1}))
        h = h // 2
    if kind == "ChtoModelsimple":
        ops.append(dict(kind="linear", name="layer6", **{"in": h, "out": h, "act": "relu"}))
        ops.append(dict(kind="linear", name="layer7", **{"in": h, "out": n_out, "act": "relu"}))
    else:
        ops.append(dict(kind="linear", name="layer6", **{"in": h, "out": h * 4, "act": "relu"}))
        ops.append(dict(kind="linear", name="layer7", **{"in": h * 4, "out": n_out, "act": "relu"}))
    ops.append(dict(kind="linear", name="layer8", **{"in": n_out, "out": n_out, "act": "none"}))
    return ops


def state_dict_shapes(kind, n_in, n_out):
    """Ordered (key, shape) list matching the reference ``state_dict()``
    (23 tensors for ChtoModelv2)."""
    out = []
    for op in chto_ops(kind, n_in, n_out):
        nm = op["name"]
        if op["kind"] == "linear":
            out.append((nm + ".weight", (op["out"], op["in"])))
            out.append((nm + ".bias", (op["out"],)))
        else:
            out.append((nm + ".layer1.weight", (op["mid"], op["in"])))
            out.append((nm + ".layer1.bias", (op["mid"],)))
            out.append((nm + ".layer2.weight", (op["out"], op["mid"])))
            out.append((nm + ".layer2.bias", (op["out"],)))
            if op["in"] != op["out"]:
                out.append((nm + ".skip_layer.weight", (op["out"], op["in"])))
    if kind == "ChtoModelv2_linear":
        out.append(("linearlayer.weight", (n_out, n_in)))
        out.append(("linearlayer.bias", (n_out,)))
    return out


def n_params(kind, n_in, n_out):
    tot = 0
    for _, shp in state_dict_shapes(kind, n_in, n_out):
        n = 1
        for s in shp:
            n *= s
        tot += n
    return tot


def macs_forward(kind, n_in, n_out):
    """Multiply-accumulates per sample through the network."""
    tot = 0
    for op in chto_ops(kind, n_in, n_out):
        if op["kind"] == "linear":
            tot += op["in"] * op["out"]
        else:
            tot += op["in"] * op["mid"] + op["mid"] * op["out"]
            if op["in"] != op["out"]:
                tot += op["in"] * op["out"]
    if kind == "ChtoModelv2_linear":
        tot += n_in * n_out
    return tot


def flops_lnl(kind, n_in, n_out):
    """Useful flops of one lnL evaluation: 2*MACs + triangular
    L^T d product + square-sum."""
    return 2 * macs_forward(kind, n_in, n_out) + n_out * (n_out + 1) + 2 * n_out


def flops_lnl_grad(kind, n_in, n_out):
    """Useful flops of lnL + d lnL/du."""
    return 4 * macs_forward(kind, n_in, n_out) + 2 * n_out * (n_out + 1)
