"""ORACLE (test infrastructure only — never imported by the product path): a small ONNX graph interpreter.

The reference runs `visual.onnx` / `text.onnx` through onnxruntime's CPU execution provider
(`/root/reference/src/onnx.rs:19-23`, `src/vision.rs:105-109`, `src/text.rs:153-162`).  onnxruntime is not installable
in this image, so this module executes the *file* — node by node, in fp32, following the public ONNX operator
specification (opset 18) — which is what `session.run` computes up to fp32 reassociation inside MatMul/Conv.  It is
the check SURVEY.md §8(c) asks for ("so the file, not the module, is what gets compared"): the functional oracle
(`oracle/reference_forward.py`) restates the architecture from parameter names, this one knows nothing about
architectures and only follows the graph.

**parity unpinned** in the sense of the task statement: it has not been run against onnxruntime itself (absent).  It
is pinned against torch (every graph in the tests was produced by `torch.onnx.export` from an `nn.Module`, and the
interpreter must reproduce that module's own fp32 output) and against an independent ONNX runtime that is in the image,
OpenCV DNN, executing the same files (`tests/test_onnx_graph_cpu.py`).

Only the operators that `torch.onnx.export` emits for the CLIP / SigLIP / ViT towers are implemented; anything else
raises `NotImplementedError` naming the operator.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_DT = {1: np.float32, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64, 2: np.uint8, 3: np.int8}


# ----------------------------------------------------------------------------- protobuf reading (independent of tools/)
def _varint(buf, pos):
    r = 0
    s = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, pos
        s += 7


def _fields(buf):
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        f, w = key >> 3, key & 7
        if w == 0:
            v, pos = _varint(buf, pos)
        elif w == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif w == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        elif w == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        else:
            raise ValueError(f"wire type {w}")
        yield f, w, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_ints(val) -> List[int]:
    out, p = [], 0
    while p < len(val):
        d, p = _varint(val, p)
        out.append(_signed(d))
    return out


def _tensor(buf, base_dir: str, mmaps: dict) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dt, name, raw, ext = 1, "", None, {}
    floats: List[float] = []
    ints: List[int] = []
    for f, w, v in _fields(buf):
        if f == 1:
            dims.extend(_packed_ints(v) if w == 2 else [_signed(v)])
        elif f == 2:
            dt = v
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
        elif f == 4:
            floats.extend(np.frombuffer(bytes(v), "<f4").tolist() if w == 2 else [struct.unpack("<f", v)[0]])
        elif f in (5, 7):
            ints.extend(_packed_ints(v) if w == 2 else [_signed(v)])
        elif f == 13:
            kv = {ff: bytes(vv).decode() for ff, _, vv in _fields(v)}
            ext[kv.get(1, "")] = kv.get(2, "")
    npdt = _DT[dt]
    if ext:
        loc = ext["location"]
        if loc not in mmaps:
            mmaps[loc] = np.memmap(os.path.join(base_dir, loc), dtype=np.uint8, mode="r")
        off = int(ext.get("offset", "0"))
        count = int(np.prod(dims)) if dims else 1
        arr = np.frombuffer(mmaps[loc], dtype=npdt, count=count, offset=off)
    elif raw is not None:
        arr = np.frombuffer(raw, dtype=npdt)
    elif floats:
        arr = np.asarray(floats, dtype=npdt)
    else:
        arr = np.asarray(ints, dtype=npdt)
    return name, np.array(arr).reshape(dims)


class Node:
    __slots__ = ("op", "inputs", "outputs", "attrs", "name")

    def __init__(self):
        self.op, self.inputs, self.outputs, self.attrs, self.name = "", [], [], {}, ""


def _attr(buf, base_dir, mmaps):
    name, val = "", None
    f_, i_, s_, t_, floats, ints = None, None, None, None, [], []
    for f, w, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            f_ = struct.unpack("<f", v)[0]
        elif f == 3:
            i_ = _signed(v)
        elif f == 4:
            s_ = bytes(v).decode()
        elif f == 5:
            t_ = _tensor(v, base_dir, mmaps)[1]
        elif f == 7:
            floats.extend(np.frombuffer(bytes(v), "<f4").tolist() if w == 2 else [struct.unpack("<f", v)[0]])
        elif f == 8:
            ints.extend(_packed_ints(v) if w == 2 else [_signed(v)])
    for cand in (t_, s_, f_, i_):
        if cand is not None:
            val = cand
            break
    if val is None:
        val = ints if ints else floats
    return name, val


class Graph:
    def __init__(self, path: str):
        base = os.path.dirname(os.path.abspath(path))
        with open(path, "rb") as f:
            buf = memoryview(f.read())
        self.initializers: Dict[str, np.ndarray] = {}
        self.nodes: List[Node] = []
        self.inputs: List[str] = []
        self.outputs: List[str] = []
        self.opset = 0
        mm: dict = {}
        for f, _, v in _fields(buf):
            if f == 7:
                for gf, _, gv in _fields(v):
                    if gf == 5:
                        n, a = _tensor(gv, base, mm)
                        self.initializers[n] = a
                    elif gf in (11, 12):
                        for vf, _, vv in _fields(gv):
                            if vf == 1:
                                (self.inputs if gf == 11 else self.outputs).append(bytes(vv).decode())
                    elif gf == 1:
                        nd = Node()
                        for nf, _, nv in _fields(gv):
                            if nf == 1:
                                nd.inputs.append(bytes(nv).decode())
                            elif nf == 2:
                                nd.outputs.append(bytes(nv).decode())
                            elif nf == 3:
                                nd.name = bytes(nv).decode()
                            elif nf == 4:
                                nd.op = bytes(nv).decode()
                            elif nf == 5:
                                k, val = _attr(nv, base, mm)
                                nd.attrs[k] = val
                        self.nodes.append(nd)
            elif f == 8:
                for of, _, ov in _fields(v):
                    if of == 2:
                        self.opset = max(self.opset, int(ov))
        self.inputs = [i for i in self.inputs if i not in self.initializers]


# ----------------------------------------------------------------------------- operators (ONNX opset 18 semantics)
def _t(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a
    a = np.asarray(a)
    return torch.from_numpy(np.array(a, order="C"))  # (np.ascontiguousarray would turn 0-d scalars into 1-d)


def _ints(t: torch.Tensor) -> List[int]:
    return [int(v) for v in t.reshape(-1).tolist()]


def _reshape(x: torch.Tensor, shape: Sequence[int], allowzero: int) -> torch.Tensor:
    shape = list(shape)
    if not allowzero:
        shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
    return x.reshape(shape)


def _slice(x, starts, ends, axes, steps):
    idx = [slice(None)] * x.dim()
    for s, e, a, st in zip(starts, ends, axes, steps):
        a = a % x.dim() if x.dim() else 0
        n = x.shape[a]
        if st > 0:
            s = max(0, min(n, s + n if s < 0 else s))
            e = max(0, min(n, e + n if e < 0 else e))
            idx[a] = slice(s, e, st)
        else:
            raise NotImplementedError("Slice with negative step")
    return x[tuple(idx)]


def _gather(x: torch.Tensor, idx: torch.Tensor, axis: int) -> torch.Tensor:
    axis %= x.dim()
    n = x.shape[axis]
    flat = idx.reshape(-1).long()
    flat = torch.where(flat < 0, flat + n, flat)
    out = x.index_select(axis, flat)
    return out.reshape(list(x.shape[:axis]) + list(idx.shape) + list(x.shape[axis + 1:]))


def _gather_nd(x: torch.Tensor, idx: torch.Tensor, batch_dims: int) -> torch.Tensor:
    if batch_dims != 0:
        raise NotImplementedError("GatherND batch_dims != 0")
    k = idx.shape[-1]
    out = x[tuple(idx[..., j].long() for j in range(k))]
    return out


_CAST = {1: torch.float32, 6: torch.int32, 7: torch.int64, 9: torch.bool, 10: torch.float16, 11: torch.float64}


def run(graph: Graph, feeds: Dict[str, np.ndarray], dtype=torch.float32) -> List[np.ndarray]:
    """Executes the graph; float initializers / inputs are computed in `dtype` (fp32 = what ORT CPU does; fp64 bounds
    the interpreter's own rounding)."""
    env: Dict[str, Optional[torch.Tensor]] = {"": None}
    for k, a in graph.initializers.items():
        t = _t(a)
        env[k] = t.to(dtype) if t.is_floating_point() else t
    for k, a in feeds.items():
        t = _t(a)
        env[k] = t.to(dtype) if t.is_floating_point() else t
    for nd in graph.nodes:
        x = [env[i] for i in nd.inputs]
        a = nd.attrs
        op = nd.op
        if op == "Constant":
            t = _t(a["value"]) if "value" in a else torch.tensor(a.get("value_float", a.get("value_int", a.get("value_ints"))))
            y = t.to(dtype) if t.is_floating_point() else t
        elif op == "Identity":
            y = x[0]
        elif op in ("Add", "Sub", "Mul", "Div", "Pow"):
            p, q = x[0], x[1]
            if op == "Add":
                y = p + q
            elif op == "Sub":
                y = p - q
            elif op == "Mul":
                y = p * q
            elif op == "Pow":
                y = torch.pow(p, q)
            elif p.is_floating_point() or q.is_floating_point():
                y = p / q
            else:
                y = torch.div(p, q, rounding_mode="trunc")
        elif op == "Mod":
            y = torch.fmod(x[0], x[1]) if a.get("fmod", 0) else torch.remainder(x[0], x[1])
        elif op == "Sqrt":
            y = torch.sqrt(x[0])
        elif op == "Erf":
            y = torch.erf(x[0])
        elif op == "Tanh":
            y = torch.tanh(x[0])
        elif op == "Sigmoid":
            y = torch.sigmoid(x[0])
        elif op == "Neg":
            y = -x[0]
        elif op == "Relu":
            y = torch.relu(x[0])
        elif op == "Softmax":
            y = torch.softmax(x[0], dim=a.get("axis", -1))
        elif op == "MatMul":
            y = torch.matmul(x[0], x[1])
        elif op == "Gemm":
            p = x[0].t() if a.get("transA", 0) else x[0]
            q = x[1].t() if a.get("transB", 0) else x[1]
            y = a.get("alpha", 1.0) * (p @ q)
            if len(x) > 2 and x[2] is not None:
                y = y + a.get("beta", 1.0) * x[2]
        elif op == "Conv":
            if len(a.get("kernel_shape", [0, 0])) != 2:
                raise NotImplementedError("Conv: only 2-D")
            pads = a.get("pads", [0, 0, 0, 0])
            if pads[0] != pads[2] or pads[1] != pads[3]:
                raise NotImplementedError("Conv: asymmetric pads")
            y = F.conv2d(x[0], x[1], x[2] if len(x) > 2 else None, stride=a.get("strides", [1, 1]),
                         padding=(pads[0], pads[1]), dilation=a.get("dilations", [1, 1]), groups=a.get("group", 1))
        elif op == "LayerNormalization":
            axis = a.get("axis", -1)
            shape = x[0].shape[axis:] if axis < 0 else x[0].shape[axis:]
            y = F.layer_norm(x[0], tuple(shape), x[1], x[2] if len(x) > 2 else None, a.get("epsilon", 1e-5))
        elif op == "BatchNormalization":   # inference form: (x - mean) / sqrt(var + eps) * scale + bias, per channel
            sh = (1, -1) + (1,) * (x[0].dim() - 2)
            y = (x[0] - x[3].reshape(sh)) / torch.sqrt(x[4].reshape(sh) + a.get("epsilon", 1e-5)) * x[1].reshape(sh) + x[2].reshape(sh)
        elif op == "GlobalAveragePool":
            y = x[0].mean(dim=(2, 3), keepdim=True)
        elif op in ("ReduceMean", "ReduceSum", "ReduceL2", "ReduceMax"):
            axes = _ints(x[1]) if len(x) > 1 and x[1] is not None else a.get("axes")
            keep = bool(a.get("keepdims", 1))
            if axes is None:
                axes = list(range(x[0].dim()))
            if op == "ReduceMean":
                y = x[0].mean(dim=axes, keepdim=keep)
            elif op == "ReduceSum":
                y = x[0].sum(dim=axes, keepdim=keep)
            elif op == "ReduceMax":
                y = x[0].amax(dim=axes, keepdim=keep)
            else:
                y = torch.sqrt((x[0] * x[0]).sum(dim=axes, keepdim=keep))
        elif op == "Clip":
            lo = x[1] if len(x) > 1 else None
            hi = x[2] if len(x) > 2 else None
            y = x[0]
            if lo is not None:
                y = torch.maximum(y, lo.to(y.dtype))
            if hi is not None:
                y = torch.minimum(y, hi.to(y.dtype))
        elif op == "ArgMax":
            y = torch.argmax(x[0], dim=a.get("axis", 0), keepdim=bool(a.get("keepdims", 1)))
        elif op == "Shape":
            s = list(x[0].shape)
            y = torch.tensor(s[a.get("start", 0):a.get("end", len(s))], dtype=torch.int64)
        elif op == "Reshape":
            y = _reshape(x[0], _ints(x[1]), a.get("allowzero", 0))
        elif op == "Flatten":
            ax = a.get("axis", 1)
            y = x[0].reshape(int(np.prod(x[0].shape[:ax])) if ax else 1, -1)
        elif op == "Transpose":
            perm = a.get("perm") or list(reversed(range(x[0].dim())))
            y = x[0].permute(perm)
        elif op == "Concat":
            y = torch.cat([t for t in x], dim=a["axis"])
        elif op == "Unsqueeze":
            axes = _ints(x[1])
            rank = x[0].dim() + len(axes)
            y = x[0]
            for ax in sorted(ax % rank for ax in axes):
                y = y.unsqueeze(ax)
        elif op == "Squeeze":
            y = x[0]
            if len(x) > 1 and x[1] is not None:
                for ax in sorted((ax % x[0].dim() for ax in _ints(x[1])), reverse=True):
                    y = y.squeeze(ax)
            else:
                y = y.squeeze()
        elif op == "Gather":
            y = _gather(x[0], x[1], a.get("axis", 0))
        elif op == "GatherND":
            y = _gather_nd(x[0], x[1], a.get("batch_dims", 0))
        elif op == "Slice":
            starts, ends = _ints(x[1]), _ints(x[2])
            axes = _ints(x[3]) if len(x) > 3 and x[3] is not None else list(range(len(starts)))
            steps = _ints(x[4]) if len(x) > 4 and x[4] is not None else [1] * len(starts)
            y = _slice(x[0], starts, ends, axes, steps)
        elif op == "Split":
            axis = a.get("axis", 0)
            if len(x) > 1 and x[1] is not None:
                parts = torch.split(x[0], _ints(x[1]), dim=axis)
            else:
                n = a.get("num_outputs", len(nd.outputs))
                parts = torch.split(x[0], -(-x[0].shape[axis] // n), dim=axis)
            for o, p in zip(nd.outputs, parts):
                env[o] = p
            continue
        elif op == "Cast":
            to = _CAST[a["to"]]
            y = x[0].to(dtype if to == torch.float32 else to)
        elif op == "Expand":
            shape = _ints(x[1])
            y = x[0] * torch.ones(shape, dtype=x[0].dtype) if x[0].dtype != torch.bool else x[0] | torch.zeros(shape, dtype=torch.bool)
        elif op == "ConstantOfShape":
            v = _t(a["value"]) if "value" in a else torch.zeros(1, dtype=torch.float32)
            v = v.to(dtype) if v.is_floating_point() else v
            y = torch.full(_ints(x[0]), v.reshape(-1)[0].item(), dtype=v.dtype)
        elif op == "Equal":
            y = x[0] == x[1]
        elif op == "Less":
            y = x[0] < x[1]
        elif op == "Greater":
            y = x[0] > x[1]
        elif op == "Not":
            y = ~x[0]
        elif op == "Where":
            y = torch.where(x[0], x[1], x[2])
        elif op == "Range":
            y = torch.arange(x[0].item(), x[1].item(), x[2].item(), dtype=x[0].dtype)
        elif op == "Trilu":
            k = int(x[1].item()) if len(x) > 1 and x[1] is not None else 0
            y = torch.triu(x[0], k) if a.get("upper", 1) else torch.tril(x[0], k)
        else:
            raise NotImplementedError(f"ONNX operator '{op}' is not implemented by the oracle interpreter")
        env[nd.outputs[0]] = y
    return [env[o].to(torch.float32).numpy() if env[o].is_floating_point() else env[o].numpy() for o in graph.outputs]


class OnnxSession:
    """The role `ort::Session` plays for the reference (`src/onnx.rs:8-46`): inputs by name, one `run`."""

    def __init__(self, path: str):
        self.graph = Graph(path)

    @property
    def input_names(self) -> List[str]:
        return list(self.graph.inputs)

    def has_input(self, name: str) -> bool:
        return name in self.graph.inputs

    @torch.no_grad()
    def run(self, feeds: Dict[str, np.ndarray], dtype=torch.float32) -> np.ndarray:
        return run(self.graph, feeds, dtype)[0]
