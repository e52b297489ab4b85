"""Protocol-buffer TEXT format, as much of it as Caffe's net / solver prototxt files use -- a small hand-written
parser and printer, so the front end needs neither protoc nor the reference's generated caffe_pb2.py (which does not
import under protobuf 6).  Reference: owl/owl/net/net_helper.py:16-27 (text_format.Merge into NetParameter /
SolverParameter).

    msg = parse(open("train_val.prototxt").read())
    for layer in msg.all("layer") + msg.all("layers"): layer.get("name"), layer.sub("convolution_param").get("num_output", 0)
"""
import re


class Msg(object):
    """One message: field name -> list of values in file order (scalars: int / float / bool / str; sub-messages: Msg)."""

    def __init__(self):
        self.fields = {}
        self.order = []          # (name, value) in file order, for printing

    def add(self, name, value):
        self.fields.setdefault(name, []).append(value)
        self.order.append((name, value))
        return self

    def all(self, name):
        return list(self.fields.get(name, []))

    def has(self, name):
        return name in self.fields

    def get(self, name, default=None):
        v = self.fields.get(name)
        return v[0] if v else default

    def sub(self, name):
        """First sub-message `name`, or an empty message (so optional blocks read with defaults)."""
        v = self.fields.get(name)
        return v[0] if v else Msg()

    def __repr__(self):
        return "Msg(%s)" % ", ".join("%s=%r" % kv for kv in self.order[:6])


_TOKEN = re.compile(r"""\s*(?:(\#[^\n]*)|("(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*')|([{}<>:;,\[\]])|([^\s{}<>:;,\[\]"'#]+))""")


def _tokens(text):
    pos, n = 0, len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return
            raise ValueError("prototxt: cannot tokenise at %r" % text[pos:pos + 30])
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            s = m.group(2)[1:-1]
            yield ("str", bytes(s, "utf-8").decode("unicode_escape") if "\\" in s else s)
        elif m.group(3) is not None:
            yield ("sym", m.group(3))
        else:
            yield ("word", m.group(4))


class Enum(str):
    """An unquoted identifier (enum value such as MAX or TRAIN): a str that prints without quotes."""


def _scalar(word):
    if word in ("true", "True"):
        return True
    if word in ("false", "False"):
        return False
    try:
        return int(word, 0)
    except ValueError:
        pass
    try:
        return float(word)
    except ValueError:
        return Enum(word)


def parse(text):
    toks = list(_tokens(text))
    i = 0

    def message(close):
        nonlocal i
        msg = Msg()
        while i < len(toks):
            kind, val = toks[i]
            if kind == "sym" and val in ("}", ">"):
                if close is None:
                    raise ValueError("prototxt: unbalanced %r" % val)
                i += 1
                return msg
            if kind == "sym" and val in (";", ","):
                i += 1
                continue
            if kind != "word":
                raise ValueError("prototxt: field name expected, got %r" % (val,))
            name = val
            i += 1
            if i < len(toks) and toks[i] == ("sym", ":"):
                i += 1
            if i >= len(toks):
                raise ValueError("prototxt: value expected after %r" % name)
            kind, val = toks[i]
            if kind == "sym" and val in ("{", "<"):
                i += 1
                msg.add(name, message("}"))
            elif kind == "sym" and val == "[":          # [a, b, c] list syntax
                i += 1
                while toks[i] != ("sym", "]"):
                    if toks[i] != ("sym", ","):
                        msg.add(name, toks[i][1] if toks[i][0] == "str" else _scalar(toks[i][1]))
                    i += 1
                i += 1
            elif kind == "str":
                # adjacent string literals concatenate
                s = val
                i += 1
                while i < len(toks) and toks[i][0] == "str":
                    s += toks[i][1]
                    i += 1
                msg.add(name, s)
            else:
                msg.add(name, _scalar(val))
                i += 1
        if close is not None:
            raise ValueError("prototxt: missing closing brace")
        return msg

    return message(None)


def dump(msg, indent=0):
    out = []
    pad = "  " * indent
    for name, v in msg.order:
        if isinstance(v, Msg):
            out.append("%s%s {\n%s%s}\n" % (pad, name, dump(v, indent + 1), pad))
        elif isinstance(v, bool):
            out.append("%s%s: %s\n" % (pad, name, "true" if v else "false"))
        elif isinstance(v, str) and not isinstance(v, Enum):
            out.append('%s%s: "%s"\n' % (pad, name, v.replace("\\", "\\\\").replace('"', '\\"')))
        elif isinstance(v, float):
            out.append("%s%s: %s\n" % (pad, name, repr(v)))
        else:
            out.append("%s%s: %s\n" % (pad, name, v))
    return "".join(out)
