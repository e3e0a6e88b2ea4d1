"""Writes tests/golden/golden_relations.json from the UNMODIFIED reference (oracle/_ref/ref_relations, built by `make -C oracle ref` where
/root/reference is mounted): getreverseindex of every position, getrightcooc / getleftcooc / getcooc of every pattern, computenpmi (right) and
computeflexgrams_fromcooc on a few corpora.  The flexgram cases are kept only where the reference's insert-while-iterating did not bite
(its result equals the clean iteration of oracle.flexgrams_fromcooc); the others are listed under "reference_diverges"."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [("hamlet", dict(t=2, l=3, Y=0.1)), ("hamlet", dict(t=2, l=5, Y=-1.0)), ("hamlet", dict(t=1, l=2, Y=0.3)),
         ("zipf4k", dict(t=2, l=4, Y=0.2)), ("zipf4k", dict(t=3, l=3, Y=0.0)), ("short", dict(t=1, l=3, Y=-1.0))]


def corpus(name):
    if name == "hamlet":
        return open(os.path.join(ROOT, "tests", "golden", "hamlet.colibri.dat"), "rb").read()[2:]
    if name == "zipf4k":
        return oracle.synth_corpus(4000, vocab=120, seed=5, mean_sentence=9).tobytes()
    return oracle.encode_corpus([[5, 6, 7, 5, 6], [5], [], [6, 7, 5, 6, 7, 8, 9], [130, 5, 6, 130]])


def main():
    out = {"generator": "tests/golden/make_golden_relations.py", "cases": []}
    for name, kw in CASES:
        body = corpus(name)
        with tempfile.NamedTemporaryFile(suffix=".colibri.dat", delete=False) as f:
            f.write(b"\xa2\x02" + body)
            path = f.name
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_relations"), "-f", path, "-t", str(kw["t"]), "-l", str(kw["l"]), "-Y", repr(kw["Y"]), "-x"],
                           capture_output=True, text=True, check=True)
        os.unlink(path)
        case = {"corpus": name, "t": kw["t"], "l": kw["l"], "threshold": kw["Y"], "G": [], "R": [], "L": [], "C": [], "O": [], "N": [], "X": []}
        for line in r.stdout.splitlines():
            p = line.split()
            if p[0] == "H":
                case["header"] = dict(x.split("=") for x in p[1:])
            elif p[0] == "G":
                case["G"].append([int(p[1]), int(p[2]), p[3:]])
            elif p[0] in "RLCO":
                case[p[0]].append([p[1], p[2], int(p[3])])
            elif p[0] == "N":
                case["N"].append([p[1], p[2], float(p[3])])
            elif p[0] == "F":
                case["F"] = dict(x.split("=") for x in p[1:])
            elif p[0] == "X":
                case["X"].append([p[1], int(p[2])])
        # computeflexgrams_fromcooc inserts into the map it iterates over (:3755-3768): flexgrams made earlier are visited again (flexgrams of
        # flexgrams) and a rehash skips or repeats patterns.  The clean iteration's result is pinned where the reference contains it with equal counts.
        pats = oracle.train(body, mintokens=kw["t"], maxlength=kw["l"], indexed=1, streamed=0).as_dict()
        _found, flex = oracle.flexgrams_fromcooc(body, pats, kw["Y"])
        ref = {bytes.fromhex(k): v for k, v in case["X"]}
        case["flex_check"] = "subset" if all(ref.get(k) == v for k, v in flex.items()) else "reference_diverges"
        out["cases"].append(case)
    with open(os.path.join(ROOT, "tests", "golden", "golden_relations.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
