"""Every `file:line` citation of the reference in this repository must point inside the cited file (round 1 shipped
src/utils.jl citations that were off by the length of another file).  The reference's line counts are a committed
fixture (tests/golden/reference_file_lengths.json, made by tests/golden/make_file_lengths.py), so the check also
runs where /root/reference does not exist."""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"((?:src|test)/[\w.]+\.jl|README\.md|Manifest\.toml|Project\.toml):(\d[\d,\- ]*\d|\d)")
SKIP = {"SURVEY.md", "VERDICT.md", "ADVICE.md", "BASELINE.md"}     # driver-written / not ours


def test_reference_citations_point_inside_the_cited_files():
    with open(os.path.join(ROOT, "tests", "golden", "reference_file_lengths.json")) as f:
        lengths = json.load(f)
    files = subprocess.run(["git", "ls-files"], cwd=ROOT, capture_output=True, text=True).stdout.split()
    assert files, "git ls-files returned nothing"
    bad, seen = [], 0
    for rel in files:
        if rel in SKIP or not rel.endswith((".py", ".c", ".h", ".cu", ".cuh", ".md", ".jl")):
            continue
        try:
            text = open(os.path.join(ROOT, rel), errors="ignore").read()
        except OSError:
            continue
        for m in PAT.finditer(text):
            name = m.group(1)
            cands = [k for k in lengths if k == name or k.endswith("/" + name)]
            if not cands:
                continue
            seen += 1
            last = max(int(x) for x in re.findall(r"\d+", m.group(2)))
            if last > max(lengths[k] for k in cands):
                bad.append(f"{rel}: {m.group(0)} (file has {max(lengths[k] for k in cands)} lines)")
    assert seen > 100, "the citation scan found suspiciously few citations"
    assert not bad, "citations past the end of the reference file:\n" + "\n".join(bad)
