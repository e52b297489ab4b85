#!/usr/bin/env python
"""Extract the golden vectors the reference's own tests hold for the hot path into JSON fixtures.

Run in the authoring container (needs /root/reference); the output files are committed so the
GPU box never reads /root/reference.  Sources:
  tests/unittest_conv_forward.cpp:7-68        conv forward, two cases (GPU tests; CPU twins DISABLED)
  tests/unittest_pooling_forward.cpp:7-81     max pooling forward, four cases
  tests/unittest_reduction.cpp:6-149          reduction closed forms on a 5x3 iota
  tests/unittest_scale.cpp:8-28               ScaleRange::Flatten
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def arrays_in(body):
    out = {}
    for m in re.finditer(r"float\s+(\w+)\[\]\s*=\s*\{([^}]*)\}", body):
        out[m.group(1)] = [float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
    return out


def scales_in(body):
    return {m.group(1): [int(v) for v in m.group(2).split(",")]
            for m in re.finditer(r"Scale\s+(\w+)\{([0-9, ]+)\}", body)}


def tests_in(path):
    src = open(path).read()
    parts = re.split(r"\nTEST\((\w+),\s*(\w+)\)\s*\{", src)
    for i in range(1, len(parts), 3):
        yield parts[i], parts[i + 1], parts[i + 2]


def conv():
    cases = []
    for suite, name, body in tests_in(os.path.join(REF, "tests/unittest_conv_forward.cpp")):
        if name.startswith("DISABLED"):
            continue
        a, s = arrays_in(body), scales_in(body)
        ci = re.search(r"ConvInfo conv_info\(([0-9, ]+)\)", body).group(1)
        ph, pw, sv, sh = [int(v) for v in ci.split(",")]
        cases.append(dict(name=name, input_size=s["input_size"], weight_size=s["weight_size"],
                          correct_size=s["correct_size"], pad_height=ph, pad_width=pw,
                          stride_vertical=sv, stride_horizontal=sh, input=a["input_raw"],
                          weight=a["weight_raw"], bias=a.get("bias_raw", [0.0] * s["weight_size"][3]),
                          # the WithPadding test compares against correct_raw + bias (line 66)
                          correct_excludes_bias="bias_raw" in a, correct=a["correct_raw"],
                          tolerance=1e-3, source="tests/unittest_conv_forward.cpp"))
    return cases


def pool():
    cases = []
    for suite, name, body in tests_in(os.path.join(REF, "tests/unittest_pooling_forward.cpp")):
        a, s = arrays_in(body), scales_in(body)
        pi = re.search(r"PoolingInfo pooling_info\(PoolingInfo::Algorithm::k(\w+),\s*([0-9, ]+)\)", body)
        nums = [int(v) for v in pi.group(2).split(",")] + [0, 0]
        h, w, sv, sh, ph, pw = nums[:6]
        cases.append(dict(name=name, algorithm=pi.group(1).lower(), input_size=s["input_size"],
                          correct_size=s["correct_size"], height=h, width=w, stride_vertical=sv,
                          stride_horizontal=sh, pad_height=ph, pad_width=pw, input=a["input_raw"],
                          correct=a["correct_raw"], tolerance=1e-3,
                          source="tests/unittest_pooling_forward.cpp"))
    return cases


def reduction():
    # unittest_reduction.cpp: 5x3 matrix filled with iota (i + 5j? see lines 9-15): value = index
    return dict(size=[5, 3], input=[float(i) for i in range(15)],
                max_dim0=[5.0 * i + 4 for i in range(3)],      # line 20
                max_dim1=[10.0 + i for i in range(5)],         # line 38
                sum_dim0=[25.0 * i + 10 for i in range(3)],    # line 56
                sum_dim1=[3.0 * i + 15 for i in range(5)],     # line 74
                source="tests/unittest_reduction.cpp:6-149")


def scale():
    # unittest_scale.cpp:10-16 (range {0,0}..{4,5}); the non-origin twin (:21-27) is the same
    # six cases shifted by one in every coordinate.
    src = open(os.path.join(REF, "tests/unittest_scale.cpp")).read()
    first = src.split("TEST(ScaleTest, FlattenIndexWithNonOriginStart)")[0]
    cases = [dict(dims=[4, 5], idx=[int(a), int(b)], flat=int(c))
             for a, b, c in re.findall(r"Flatten\(\{(\d+), (\d+)\}\), (\d+)\)", first)]
    assert len(cases) == 6
    return dict(cases=cases, source="tests/unittest_scale.cpp:8-28")


if __name__ == "__main__":
    json.dump(conv(), open(os.path.join(HERE, "conv_forward.json"), "w"))
    json.dump(pool(), open(os.path.join(HERE, "pooling_forward.json"), "w"))
    json.dump(reduction(), open(os.path.join(HERE, "reduction.json"), "w"))
    json.dump(scale(), open(os.path.join(HERE, "scale_flatten.json"), "w"))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".json")))
