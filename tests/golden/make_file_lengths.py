"""Line counts of the reference's files (run in the authoring container, where /root/reference exists): the fixture
that tests/test_citations.py checks every `file:line` citation of this repository against."""
import json
import os

REF = "/root/reference"
out = {}
for root, _, files in os.walk(REF):
    if "/.git" in root:
        continue
    for f in files:
        if f.endswith((".jl", ".md", ".toml")):
            p = os.path.join(root, f)
            with open(p, errors="ignore") as fh:
                out[os.path.relpath(p, REF)] = sum(1 for _ in fh)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_file_lengths.json"), "w") as fh:
    json.dump(dict(sorted(out.items())), fh, indent=1)
print(len(out), "files")
