"""Writes tests/golden/golden_stats.json from the UNMODIFIED reference (oracle/_ref/ref_relations -S): totaloccurrencesingroup /
totalpatternsingroup / totalwordtypesingroup (include/patternmodel.h:1903-2030) of indexed models for every (category, n).
Only cases trained with a threshold >= 2: with MINTOKENS = 1 the reference's train() asks for totalwordtypesingroup(NGRAM, 1) on the way
(:1205), which fills its statistics cache with that one group and makes every later question about another group answer 0 (:2024-2030)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_relations import corpus  # noqa: E402

CASES = [("hamlet", dict(t=2, l=4, s=1)), ("hamlet", dict(t=2, l=6, s=1)), ("hamlet", dict(t=2, l=3, s=0)), ("zipf4k", dict(t=2, l=4, s=1)), ("zipf4k", dict(t=3, l=5, s=0))]


def main():
    import tempfile

    out = {"generator": "tests/golden/make_golden_stats.py", "cases": []}
    for name, kw in CASES:
        with tempfile.NamedTemporaryFile(suffix=".colibri.dat", delete=False) as f:
            f.write(b"\xa2\x02" + corpus(name))
            path = f.name
        args = [os.path.join(ROOT, "oracle", "_ref", "ref_relations"), "-f", path, "-t", str(kw["t"]), "-l", str(kw["l"]), "-S"] + (["-s"] if kw["s"] else [])
        r = subprocess.run(args, capture_output=True, text=True, check=True)
        os.unlink(path)
        rows = [[int(x) for x in line.split()[1:]] for line in r.stdout.splitlines() if line.startswith("S ")]
        out["cases"].append({"corpus": name, "t": kw["t"], "l": kw["l"], "skipgrams": kw["s"], "S": rows})
    with open(os.path.join(ROOT, "tests", "golden", "golden_stats.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
