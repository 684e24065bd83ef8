"""Minimal ONNX protobuf wire-format writer/reader (no `onnx` package in this image).

Field numbers follow the public onnx.proto3 schema (also listed in SURVEY.md Appendix B):
  ModelProto   1 ir_version, 2 producer_name, 3 producer_version, 7 graph, 8 opset_import, 14 metadata_props
  GraphProto   1 node, 2 name, 5 initializer, 11 input, 12 output
  NodeProto    1 input, 2 output, 3 name, 4 op_type, 5 attribute
  TensorProto  1 dims, 2 data_type, 8 name, 9 raw_data, 13 external_data, 14 data_location
  ValueInfoProto 1 name, 2 type;  TypeProto 1 tensor_type{1 elem_type, 2 shape{1 dim{1 dim_value,2 dim_param}}}

The reference's exporter (`/root/reference/pull_onnx.py:159-181`) produces `visual.onnx` / `text.onnx` with one
input, one output, a dynamic batch axis, opset 18, and weights in a sibling `*.onnx.data` file
(`/root/reference/src/model_manager.rs:16-17` requires both files).  This module writes files with that I/O
contract and that external-data layout, and reads them back (used by the CPU oracle and the tests).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

FLOAT, INT32, INT64, FLOAT16, BFLOAT16 = 1, 6, 7, 10, 16
_NP2ONNX = {np.dtype(np.float32): FLOAT, np.dtype(np.int64): INT64, np.dtype(np.int32): INT32,
            np.dtype(np.float16): FLOAT16}
_ONNX2NP = {FLOAT: np.float32, INT64: np.int64, INT32: np.int32, FLOAT16: np.float16}


# ----------------------------------------------------------------------------- encoding primitives
def _varint(n: int) -> bytes:
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def f_str(field: int, s: str) -> bytes:
    return f_bytes(field, s.encode("utf-8"))


def f_float(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", v)


# ----------------------------------------------------------------------------- messages
def string_string(key: str, value: str) -> bytes:
    return f_str(1, key) + f_str(2, value)


def tensor_proto(name: str, arr: Optional[np.ndarray] = None, *, dims: Optional[Iterable[int]] = None,
                 data_type: Optional[int] = None, external: Optional[Tuple[str, int, int]] = None) -> bytes:
    """Inline (`raw_data`) when `arr` is given and `external` is None; otherwise an external-data record
    (location, offset, length) relative to the .onnx file's directory."""
    if arr is not None:
        dims = arr.shape
        data_type = _NP2ONNX[arr.dtype]
    out = bytearray()
    packed = b"".join(_varint(int(d)) for d in dims)
    out += f_bytes(1, packed)
    out += f_varint(2, int(data_type))
    out += f_str(8, name)
    if external is None:
        out += f_bytes(9, np.ascontiguousarray(arr).tobytes())
    else:
        loc, off, length = external
        out += f_bytes(13, string_string("location", loc))
        out += f_bytes(13, string_string("offset", str(off)))
        out += f_bytes(13, string_string("length", str(length)))
        out += f_varint(14, 1)  # data_location = EXTERNAL
    return bytes(out)


def value_info(name: str, elem_type: int, shape: Iterable) -> bytes:
    dims = b""
    for d in shape:
        if isinstance(d, str):
            dims += f_bytes(1, f_str(2, d))
        else:
            dims += f_bytes(1, f_varint(1, int(d)))
    tensor_type = f_varint(1, elem_type) + f_bytes(2, dims)
    return f_str(1, name) + f_bytes(2, f_bytes(1, tensor_type))


def attr_int(name: str, v: int) -> bytes:
    return f_str(1, name) + f_varint(3, v) + f_varint(20, 2)


def attr_float(name: str, v: float) -> bytes:
    return f_str(1, name) + f_float(2, v) + f_varint(20, 1)


def attr_ints(name: str, vs: Iterable[int]) -> bytes:
    return f_str(1, name) + b"".join(f_varint(8, int(v)) for v in vs) + f_varint(20, 7)


def attr_str(name: str, s: str) -> bytes:
    return f_str(1, name) + f_bytes(4, s.encode()) + f_varint(20, 3)


def attr_tensor(name: str, t: bytes) -> bytes:
    return f_str(1, name) + f_bytes(5, t) + f_varint(20, 4)


def node(op_type: str, inputs: List[str], outputs: List[str], name: str = "", attrs: Iterable[bytes] = ()) -> bytes:
    out = bytearray()
    for i in inputs:
        out += f_str(1, i)
    for o in outputs:
        out += f_str(2, o)
    if name:
        out += f_str(3, name)
    out += f_str(4, op_type)
    for a in attrs:
        out += f_bytes(5, a)
    return bytes(out)


class ModelWriter:
    """Streams a ModelProto to `<path>` with every tensor >= `external_threshold` bytes stored in `<path>.data`."""

    def __init__(self, path: str, graph_name: str, producer: str = "clipb200-export-synthetic",
                 opset: int = 18, external_threshold: int = 1024):
        self.path = path
        self.data_name = os.path.basename(path) + ".data"
        self.graph_name = graph_name
        self.producer = producer
        self.opset = opset
        self.external_threshold = external_threshold
        self._data = open(path + ".data", "wb")
        self._offset = 0
        self._inits: List[bytes] = []
        self._nodes: List[bytes] = []
        self._inputs: List[bytes] = []
        self._outputs: List[bytes] = []
        self._meta: List[bytes] = []

    def add_initializer(self, name: str, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr)
        nbytes = arr.nbytes
        if nbytes >= self.external_threshold:
            pad = (-self._offset) % 64  # keep external tensors 64-byte aligned like torch's exporter does
            if pad:
                self._data.write(b"\0" * pad)
                self._offset += pad
            self._data.write(arr.tobytes() if nbytes < (1 << 28) else memoryview(arr).cast("B"))
            self._inits.append(tensor_proto(name, dims=arr.shape, data_type=_NP2ONNX[arr.dtype],
                                            external=(self.data_name, self._offset, nbytes)))
            self._offset += nbytes
        else:
            self._inits.append(tensor_proto(name, arr))

    def add_node(self, n: bytes) -> None:
        self._nodes.append(n)

    def add_input(self, name: str, elem_type: int, shape: Iterable) -> None:
        self._inputs.append(value_info(name, elem_type, shape))

    def add_output(self, name: str, elem_type: int, shape: Iterable) -> None:
        self._outputs.append(value_info(name, elem_type, shape))

    def add_metadata(self, key: str, value: str) -> None:
        self._meta.append(string_string(key, value))

    def close(self) -> None:
        self._data.close()
        g = bytearray()
        for n in self._nodes:
            g += f_bytes(1, n)
        g += f_str(2, self.graph_name)
        for t in self._inits:
            g += f_bytes(5, t)
        for i in self._inputs:
            g += f_bytes(11, i)
        for o in self._outputs:
            g += f_bytes(12, o)
        m = bytearray()
        m += f_varint(1, 8)  # ir_version 8
        m += f_str(2, self.producer)
        m += f_str(3, "1")
        m += f_bytes(7, bytes(g))
        m += f_bytes(8, f_str(1, "") + f_varint(2, self.opset))
        for kv in self._meta:
            m += f_bytes(14, kv)
        with open(self.path, "wb") as f:
            f.write(m)


# ----------------------------------------------------------------------------- decoding
def _read_varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _fields(buf: memoryview):
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wire, v
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            yield field, wire, buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            yield field, wire, bytes(buf[pos:pos + 4])
            pos += 4
        elif wire == 1:
            yield field, wire, bytes(buf[pos:pos + 8])
            pos += 8
        else:
            raise ValueError(f"unsupported wire type {wire}")


def _parse_kv(buf: memoryview) -> Tuple[str, str]:
    k = v = ""
    for f, _, val in _fields(buf):
        if f == 1:
            k = bytes(val).decode()
        elif f == 2:
            v = bytes(val).decode()
    return k, v


def _parse_tensor(buf: memoryview, base_dir: str, mmaps: Dict[str, np.memmap]) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = FLOAT
    name = ""
    raw = None
    ext: Dict[str, str] = {}
    float_data: List[float] = []
    int64_data: List[int] = []
    for f, w, val in _fields(buf):
        if f == 1:
            if w == 2:
                p = 0
                while p < len(val):
                    d, p = _read_varint(val, p)
                    dims.append(d)
            else:
                dims.append(val)
        elif f == 2:
            dtype = val
        elif f == 8:
            name = bytes(val).decode()
        elif f == 9:
            raw = val
        elif f == 13:
            k, v = _parse_kv(val)
            ext[k] = v
        elif f == 4:
            if w == 2:
                float_data.extend(np.frombuffer(bytes(val), dtype="<f4").tolist())
            else:
                float_data.append(struct.unpack("<f", val)[0])
        elif f == 7:
            if w == 2:
                p = 0
                while p < len(val):
                    d, p = _read_varint(val, p)
                    int64_data.append(d - (1 << 64) if d >= (1 << 63) else d)
            else:
                int64_data.append(val)
    np_dtype = _ONNX2NP[dtype]
    if ext:
        loc = ext["location"]
        if loc not in mmaps:
            mmaps[loc] = np.memmap(os.path.join(base_dir, loc), dtype=np.uint8, mode="r")
        off = int(ext.get("offset", "0"))
        count = int(np.prod(dims)) if dims else 1
        length = int(ext.get("length", str(count * np.dtype(np_dtype).itemsize)))
        arr = np.frombuffer(mmaps[loc], dtype=np_dtype, count=length // np.dtype(np_dtype).itemsize, offset=off)
    elif raw is not None:
        arr = np.frombuffer(bytes(raw), dtype=np_dtype)
    elif float_data:
        arr = np.asarray(float_data, dtype=np.float32)
    else:
        arr = np.asarray(int64_data, dtype=np_dtype)
    return name, arr.reshape(dims)


def read_model(path: str) -> dict:
    """Returns {"initializers": {name: ndarray}, "inputs": [names], "outputs": [names], "metadata": {k: v},
    "nodes": [(op_type, inputs, outputs)]}."""
    base_dir = os.path.dirname(os.path.abspath(path))
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out = {"initializers": {}, "inputs": [], "outputs": [], "metadata": {}, "nodes": []}
    mmaps: Dict[str, np.memmap] = {}
    for f, _, val in _fields(buf):
        if f == 7:
            for gf, _, gval in _fields(val):
                if gf == 5:
                    name, arr = _parse_tensor(gval, base_dir, mmaps)
                    out["initializers"][name] = arr
                elif gf in (11, 12):
                    for vf, _, vval in _fields(gval):
                        if vf == 1:
                            out["inputs" if gf == 11 else "outputs"].append(bytes(vval).decode())
                elif gf == 1:
                    op, ins, outs = "", [], []
                    for nf, _, nval in _fields(gval):
                        if nf == 1:
                            ins.append(bytes(nval).decode())
                        elif nf == 2:
                            outs.append(bytes(nval).decode())
                        elif nf == 4:
                            op = bytes(nval).decode()
                    out["nodes"].append((op, ins, outs))
        elif f == 14:
            k, v = _parse_kv(val)
            out["metadata"][k] = v
    return out
