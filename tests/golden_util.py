import gzip
import json
from functools import lru_cache
from pathlib import Path

GOLDEN = Path(__file__).resolve().parent / "golden" / "golden_v1.json.gz"


@lru_cache(maxsize=1)
def load():
    with gzip.open(GOLDEN, "rt") as f:
        return json.load(f)


def case(name):
    for c in load()["cases"]:
        if c["name"] == name:
            return c
    raise KeyError(name)


CASE_NAMES = ["c1_demo_circular", "c2_small_circular", "c2_small_circular_scoring2",
              "c2_small_circular_homopolymers", "c4_small_linear",
              "fresh_31_topology0", "fresh_32_topology0", "fresh_33_topology1",
              "c1_demo_linear", "c2_fullsize_circular", "iupac_maps_circular", "c2_small_fragments",
              "c2_small_circular_scoring3", "c4_small_linear_homopolymers"]
