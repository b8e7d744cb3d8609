"""Golden vectors for the OCR-quality harness, generated FROM THE UNMODIFIED REFERENCE: ``calculate_cer`` and
``get_ground_truth_from_filename`` of /root/reference/evaluation/eval.py:22-33 and ``sort_license_plate_detections`` of
/root/reference/my_utils/utils.py:7-72.  The function definitions are executed from the reference's own source text (eval.py's imports --
loguru, tqdm, the YOLOv5 tree, the ``Levenshtein`` package, which is absent from this image -- are not needed); ``Levenshtein.distance``
is stood in for by a memoised recursive edit distance written here from the textbook definition (unit costs), independent of the
iterative implementation in lpsr_b200/evaluation.py.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_eval.py
"""
import ast
import functools
import json
import os
import random
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def functions(src, names, ns):
    tree = ast.parse(open(src).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), src, "exec"), ns)
    return [ns[n] for n in names]


def edit_distance(a, b):
    @functools.lru_cache(maxsize=None)
    def d(i, j):
        if i == 0:
            return j
        if j == 0:
            return i
        return min(d(i - 1, j) + 1, d(i, j - 1) + 1, d(i - 1, j - 1) + (a[i - 1] != b[j - 1]))
    return d(len(a), len(b))


def main():
    lev = types.SimpleNamespace(distance=edit_distance)
    cer, gt_of = functions("/root/reference/evaluation/eval.py", ["calculate_cer", "get_ground_truth_from_filename"], {"os": os, "Levenshtein": lev})
    (sort_det,) = functions("/root/reference/my_utils/utils.py", ["sort_license_plate_detections"], {})
    rng = random.Random(0)
    alphabet = "ABCDEFGHKLMNPSTUVXYZ0123456789"
    pairs = [("", ""), ("", "A"), ("A", ""), ("51F12345", "51F12345"), ("51F12345", "51F1234"), ("51F12345", "5IF12B45"), ("KITTEN", "SITTING"),
             ("SATURDAY", "SUNDAY"), ("FLAW", "LAWN"), ("30A99999", "99999A03")]
    for _ in range(60):
        g = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 10)))
        o = list(g)
        for _ in range(rng.randint(0, 4)):
            op = rng.randint(0, 2)
            pos = rng.randint(0, max(len(o), 1) - 1) if o else 0
            if op == 0 and o:
                del o[pos]
            elif op == 1:
                o.insert(pos, rng.choice(alphabet))
            elif o:
                o[pos] = rng.choice(alphabet)
        pairs.append((g, "".join(o)))
    cer_cases = [{"gt": g, "ocr": o, "distance": edit_distance(g, o), "cer": cer(g, o)} for g, o in pairs]
    names = ["51f-123.45.jpg", "30A99999.PNG", "dir/sub/29h1.2345.jpeg", "noext", "a.b.c.png"]
    name_cases = [{"name": n, "gt": gt_of(os.path.basename(n))} for n in names]
    sort_cases = []
    for k in range(40):
        n = rng.randint(0, 12)
        two_rows = rng.random() < 0.5
        dets = []
        for i in range(n):
            row = (i % 2) if two_rows else 0
            x1 = rng.uniform(0, 100); y1 = row * rng.uniform(8, 40) + rng.uniform(0, 6)
            w = rng.uniform(4, 12); hgt = rng.uniform(8, 20)
            if rng.random() < 0.2:   # integer boxes (what bb_scale=True returns) incl. exact ties
                x1, y1, w, hgt = float(int(x1)), float(int(y1)), float(int(w)), float(int(hgt))
            dets.append([rng.choice(alphabet), round(rng.random(), 3), [x1, y1, x1 + w, y1 + hgt]])
        tagged = [[c, p, tuple(b)] for c, p, b in dets]
        out = sort_det([list(t) for t in tagged])
        # identify outputs by identity of (class, conf, bbox) -- duplicates are disambiguated by first unused match
        used, order = set(), []
        for o in out:
            for i, t in enumerate(tagged):
                if i not in used and t[0] == o[0] and t[1] == o[1] and tuple(t[2]) == tuple(o[2]):
                    used.add(i); order.append(i); break
        sort_cases.append({"detections": dets, "order": order, "text": "".join(o[0].upper() for o in out)})
    json.dump({"cer": cer_cases, "names": name_cases, "sort": sort_cases}, open(os.path.join(HERE, "eval_cases.json"), "w"), indent=0)
    print(len(cer_cases), "cer cases,", len(name_cases), "names,", len(sort_cases), "sort cases")


if __name__ == "__main__":
    main()
