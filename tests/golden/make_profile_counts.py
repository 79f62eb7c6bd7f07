#!/usr/bin/env python
"""Extracts the call counts of the reference's own profile (/root/reference/output.svg, a gprof2dot graph of one
iteration over the 157 LEDs of dataset_dogStomach.json) into tests/golden/profile_call_counts.json.

These edges `runFPM -> <cvComplex helper>  N x` are the only artefact the reference holds about the arithmetic of
the hot path: they pin how often every un-vendored cvComplex helper runs per LED, i.e.
the op multiset that oracle/cv2_mirror.py replays (9 complexMultiply, 4 complexAbs, 3 complexDivide, 2 complexConj,
5 fftShift, 1 ifft2, 1 fft2 per LED; +1 complexMultiply, +3 fftShift, +1 fft2 at initialisation).
Runs ONLY in the authoring container (needs /root/reference); the JSON it writes is committed.
"""
import html
import json
import os
import re

REF = "/root/reference/output.svg"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_call_counts.json")


def main():
    lines = open(REF).read().split("\n")
    edges = {}
    for i, ln in enumerate(lines):
        m = re.search(r'class="edge"><title>(.*)</title>', ln)
        if not m:
            continue
        title = html.unescape(m.group(1))
        if "->" not in title:
            continue
        src, dst = title.split("->", 1)
        cnt = None
        for nxt in lines[i + 1:i + 8]:
            c = re.search(r">(\d+)\N{MULTIPLICATION SIGN}</text>", nxt)
            if c:
                cnt = int(c.group(1))
                break
        if cnt is not None:
            edges[(src.strip(), dst.strip())] = (cnt, i + 1)
    out = {"source": "output.svg of Xiongda337/fpm-OpenCV (edge labels `N x`)", "runFPM": {}, "other": {}}
    for (src, dst), (cnt, line) in sorted(edges.items()):
        name = dst.split("(")[0]
        if src.startswith("runFPM(") and name in ("fftShift", "fft2", "ifft2", "complexMultiply", "complexDivide", "complexAbs", "complexConj"):
            out["runFPM"][name] = {"calls": cnt, "svg_line": line}
    # cv::dft node: total calls and the number of 1-D DFT_64f transforms below it (K2)
    for i, ln in enumerate(lines):
        t = re.search(r'class="node"><title>(.*)</title>', ln)
        if t and html.unescape(t.group(1)).startswith("cv::DFT_64f"):
            for nxt in lines[i + 1:i + 12]:
                c = re.search(r">(\d+)\N{MULTIPLICATION SIGN}</text>", nxt)
                if c:
                    out["other"]["cv::DFT_64f"] = {"calls": int(c.group(1)), "svg_line": i + 1}
                    break
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
