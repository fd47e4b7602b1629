"""Writes tests/golden/reference_known_answers.json.

The reference (darchr/EmbeddingTables.jl) is Julia and cannot execute in this image, so these
fixtures are not produced by running it: they are the known-answer vectors and worked examples
the reference itself holds, transcribed literally with their file:line.  They are every
golden vector the reference has for this path.

Run:  python tests/golden/make_golden.py
"""
import json
import os

G = {}

# test/misc.jl:9-10 -- traversal order of `columns`
G["columns"] = {
    "source": "test/misc.jl:3-10",
    "vector": {"x": [1, 2, 3, 4], "expect": [[1, 1], [2, 2], [3, 3], [4, 4]]},
    # y = [1 2; 3 4]  (row-major literal; stored here as rows)
    "matrix": {"rows": [[1, 2], [3, 4]], "expect": [[1, 1], [1, 3], [2, 2], [2, 4]]},
}

# test/misc.jl:39-71 -- histogram! known answer
G["histogram"] = {
    "source": "test/misc.jl:33-72",
    "A": [2] * 10 + [1] * 5 + [20] * 3 + [5],
    "maxindex": 20,
    # key -> [order, count]
    "expect": {"2": [1, 10], "1": [2, 5], "20": [3, 3], "5": [4, 1]},
    "keys_in_insertion_order": [2, 1, 20, 5],
}

# test/misc.jl:79-107 -- index! known answer (both Sparse and Dense indexers, run twice)
A = [10, 4, 10, 100, 4, 4, 4, 1, 9, 10, 5]
G["index"] = {
    "source": "test/misc.jl:74-110",
    "A": A,
    "maxindex": max(A),
    "cumulative": [[10, 1], [4, 4], [100, 8], [1, 9], [9, 10], [5, 11], [0, 12]],
    "map": [1, 3, 10, 2, 5, 6, 7, 4, 8, 9, 11],
}

# README.md:32-73 -- integer table lookups
G["readme_lookup"] = {
    "source": "README.md:32-73",
    "data_rows": [[1, 2, 3, 4, 5]] * 5,
    "gather": {"inds": [1, 3, 4, 4, 2, 5], "expect_rows": [[1, 3, 4, 4, 2, 5]] * 5},
    "pooled": {"inds_rows": [[1, 4], [2, 5]], "expect_rows": [[3, 9]] * 5},
}

# README.md:113-160 -- maplookup over two integer tables
G["readme_maplookup"] = {
    "source": "README.md:113-160",
    "A_rows": [[1, 2, 3]] * 2,
    "B_rows": [[10, 20, 30]] * 2,
    "iA": [1, 2, 1],
    "iB": [2, 1, 1],
    "expect_A_rows": [[1, 2, 1]] * 2,
    "expect_B_rows": [[20, 10, 10]] * 2,
}

# README.md:190-232 -- pullback + Descent(0.1) on a zero 4x4 table
G["readme_update"] = {
    "source": "README.md:190-232",
    "table_shape": [4, 4],
    "inds": [1, 3, 4],
    "adjoint_rows": [[1, 5, 9], [2, 6, 10], [3, 7, 11], [4, 8, 12]],
    "eta": 0.1,
    "expect_rows": [[-0.1, 0.0, -0.5, -0.9], [-0.2, 0.0, -0.6, -1.0], [-0.3, 0.0, -0.7, -1.1],
                    [-0.4, 0.0, -0.8, -1.2]],
    "note": "printed with Julia's shortest round-trip Float32 formatting; compare as float32(x)",
}

# test/constructors.jl:5-25 -- constructor validation
G["constructors"] = {
    "source": "test/constructors.jl:1-25",
    "ok": [["static", 64, [64, 10]], ["dynamic", None, [65, 10]], ["static", 65, [65, 10]]],
    "argument_error": [["static", 32, [64, 10]], ["static", 64.0, [64, 10]]],
}

here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "reference_known_answers.json"), "w") as f:
    json.dump(G, f, indent=1, sort_keys=True)
print("wrote", os.path.join(here, "reference_known_answers.json"))
