"""Protocol-buffer WIRE format, as much of it as a .caffemodel needs: NetParameter -> layers -> (name, type, blobs).
Hand-written so the converter needs neither protoc nor the reference's generated caffe_pb2.py.  Field numbers are Caffe's
(reference copy of the schema: owl/owl/net/caffe/caffe.proto): NetParameter.layers = 2 (V1LayerParameter: name 4, type 5
enum, blobs 6), NetParameter.layer = 100 (LayerParameter: name 1, type 2 string, blobs 7); BlobProto num/channels/height/
width = 1..4, data = 5 (packed floats), shape = 7 (BlobShape.dim = 1, packed int64).  `write_caffemodel` produces such a
file (tests, and exporting weights back to Caffe)."""
import struct

import numpy as np

V1_TYPES = {4: "Convolution", 14: "InnerProduct"}      # V1LayerParameter.LayerType values the converter needs


def _varint(buf, pos):
    r = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def _fields(buf):
    """-> [(field number, wire type, value)]: varints as int, 32/64-bit as raw bytes, length-delimited as memoryview."""
    pos, n, out = 0, len(buf), []
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 5:
            v, pos = bytes(buf[pos:pos + 4]), pos + 4
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        else:
            raise ValueError("caffemodel: unsupported wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def _blob(buf):
    dims, legacy, chunks = [], {}, []
    for fn, wt, v in _fields(buf):
        if fn in (1, 2, 3, 4) and wt == 0:
            legacy[fn] = v
        elif fn == 5:
            chunks.append(np.frombuffer(bytes(v), "<f4") if wt == 2 else np.frombuffer(v, "<f4"))
        elif fn == 7 and wt == 2:
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    p = 0
                    vb = bytes(v2)
                    while p < len(vb):
                        d, p = _varint(vb, p)
                        dims.append(d)
                elif f2 == 1:
                    dims.append(v2)
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
    if not dims and legacy:
        dims = [legacy.get(i, 1) for i in (1, 2, 3, 4)]
    return {"shape": dims, "data": data.astype(np.float32)}


def read_caffemodel(path):
    """-> [{"name", "type", "blobs": [{"shape", "data"}]}] for every layer of the file, old (V1) or new format."""
    buf = memoryview(open(path, "rb").read())
    layers = []
    for fn, wt, v in _fields(buf):
        if wt != 2 or fn not in (2, 100):
            continue
        v1 = fn == 2
        name, ltype, blobs = "", "", []
        for f2, w2, v2 in _fields(v):
            if f2 == (4 if v1 else 1) and w2 == 2:
                name = bytes(v2).decode()
            elif v1 and f2 == 5 and w2 == 0:
                ltype = V1_TYPES.get(v2, "V1:%d" % v2)
            elif not v1 and f2 == 2 and w2 == 2:
                ltype = bytes(v2).decode()
            elif f2 == (6 if v1 else 7) and w2 == 2:
                blobs.append(_blob(v2))
        layers.append({"name": name, "type": ltype, "blobs": blobs})
    return layers


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def write_caffemodel(path, layers, v1=False):
    """layers: [{"name", "type" ("Convolution" | "InnerProduct" | ...), "blobs": [np.ndarray (shape = Caffe's)]}]."""
    rev = {v: k for k, v in V1_TYPES.items()}
    out = bytearray()
    for l in layers:
        body = bytearray()
        if v1:
            body += _ld(4, l["name"].encode()) + _enc_varint((5 << 3) | 0) + _enc_varint(rev.get(l["type"], 0))
        else:
            body += _ld(1, l["name"].encode()) + _ld(2, l["type"].encode())
        for b in l["blobs"]:
            b = np.asarray(b, np.float32)
            blob = bytearray()
            if v1:
                dims = ([1] * (4 - b.ndim) + list(b.shape))[:4]
                for i, d in enumerate(dims):
                    blob += _enc_varint(((i + 1) << 3) | 0) + _enc_varint(int(d))
            else:
                blob += _ld(7, _ld(1, b"".join(_enc_varint(int(d)) for d in b.shape)))
            blob += _ld(5, b.astype("<f4").tobytes())
            body += _ld(6 if v1 else 7, bytes(blob))
        out += _ld(2 if v1 else 100, bytes(body))
    with open(path, "wb") as f:
        f.write(bytes(out))
