#!/usr/bin/env python
"""Golden vectors produced by THE REFERENCE'S OWN CODE: the Python twin of the hot path,
/root/reference/src/main/python/variants_pca.py (`prepare_call_data` :19-52, `calculate_similarity_matrix` :54-82,
`center_matrix` :84-121 -- the encode, similarity and centering steps of VariantsPca.scala:153-223 restated by the
reference's authors).  The Scala driver cannot run here (no JVM), and the twin's last step (`perform_pca`, :123-152) needs the
JVM through py4j; but its first three functions are pure Python over RDD operations, so they are EXECUTED here:

  * the file is read from /root/reference at generation time (never copied into this repository), and the Python-2-only
    constructs in it are rewritten mechanically -- the substitutions `2to3` would make (fix_tuple_params, fix_xrange, fix_print,
    fix_dict), listed in REWRITES below, each asserted to match exactly the expected number of times;
  * `pyspark` is replaced by an in-memory stand-in that implements the dozen RDD methods the three functions call (map, filter,
    mapPartitions, reduceByKey, groupByKey, sortByKey, cache, collect, broadcast) with Spark's semantics for them;
  * `numpy.int` (removed from numpy 1.24) is pointed at `int`, which is what it was an alias of.

Outputs (tests/golden/reference_twin.json): for each case the input variant records and id -> index map, and what the
reference's functions returned: the call rows, all N^2 ((y, x), count) similarity records, and the centered rows.  The CPU tests
(tests/test_reference_twin.py) hold the oracle to these vectors bit for bit; the GPU tests hold the CUDA path to them.

One documented divergence inside the reference itself: the twin keeps a call when `any(c['genotype'])` (:36), i.e. it counts a
no-call (-1) as variation; the Scala driver uses `_ > 0` (VariantsPca.scala:58), which this repository follows.  The cases below
therefore contain no negative allele, where the two rules agree; `case_nocall_divergence` records what the twin does with one.

    python tests/golden/make_reference_twin_golden.py        # needs /root/reference
"""
import json
import re
import sys
import types
from pathlib import Path

import numpy

REF = Path("/root/reference/src/main/python/variants_pca.py")
OUT = Path(__file__).resolve().parent / "reference_twin.json"

# (regex, replacement, expected number of matches): the Python 2 -> 3 rewrites of the twin's source
REWRITES = [
    (r"lambda \(\(y, x\), v\): \(y, \(x, float\(v\)\)\)", "lambda yx_v: (yx_v[0][0], (yx_v[0][1], float(yx_v[1])))", 1),   # fix_tuple_params
    (r"lambda \(y, xvs\): sum\(v for \(x, v\) in xvs\)", "lambda y_xvs: sum(v for (x, v) in y_xvs[1])", 1),              # fix_tuple_params
    (r"def center_rows\(\(row, col_vals\)\):", "def center_rows(row_col_vals):\n        row, col_vals = row_col_vals", 1),    # fix_tuple_params
    (r"\bxrange\(", "range(", 2),                                                                                          # fix_xrange
    (r"\.iteritems\(\)", ".items()", 1),                                                                                   # fix_dict
    (r"print '%s\\t%s' % \(name, '\\t'\.join\(str\(c\) for c in components\)\)",
     "print('%s\\t%s' % (name, '\\t'.join(str(c) for c in components)))", 1),                                              # fix_print
]


# ------------------------------------------------------------------------------------------------ pyspark stand-in
class Broadcast:
    def __init__(self, value):
        self.value = value


class LocalRDD:
    """The RDD operations the twin calls, with Spark's semantics, on a list of partitions (lists)."""

    def __init__(self, partitions):
        self.partitions = [list(p) for p in partitions]

    def map(self, f):
        return LocalRDD([[f(x) for x in p] for p in self.partitions])

    def filter(self, f):
        return LocalRDD([[x for x in p if f(x)] for p in self.partitions])

    def mapPartitions(self, f):
        return LocalRDD([list(f(iter(p))) for p in self.partitions])

    def reduceByKey(self, op):
        acc = {}
        for p in self.partitions:
            for k, v in p:
                acc[k] = op(acc[k], v) if k in acc else v
        return LocalRDD([list(acc.items())])

    def groupByKey(self):
        acc = {}
        for p in self.partitions:
            for k, v in p:
                acc.setdefault(k, []).append(v)
        return LocalRDD([list(acc.items())])

    def sortByKey(self, ascending=True):
        items = sorted((kv for p in self.partitions for kv in p), key=lambda kv: kv[0], reverse=not ascending)
        return LocalRDD([items])

    def cache(self):
        return self

    def collect(self):
        return [x for p in self.partitions for x in p]


class SparkContext:
    _active_spark_context = None

    def __init__(self, conf=None):
        SparkContext._active_spark_context = self

    def broadcast(self, value):
        return Broadcast(value)

    def parallelize(self, data, num_slices=1):
        data = list(data)
        per = (len(data) + num_slices - 1) // max(1, num_slices)
        return LocalRDD([data[i:i + per] for i in range(0, max(len(data), 1), max(per, 1))])

    def __getattr__(self, name):      # _jvm, _jsc: the JVM side does not exist here
        raise AttributeError("no JVM behind this stand-in: " + name)


def install_pyspark_stub():
    ps = types.ModuleType("pyspark")
    ps.SparkContext = SparkContext
    conf = types.ModuleType("pyspark.conf")
    conf.SparkConf = lambda: None
    ps.conf = conf
    ps.serializers = types.ModuleType("pyspark.serializers")
    ps.rdd = types.ModuleType("pyspark.rdd")
    mllib = types.ModuleType("pyspark.mllib")
    mllib.common = types.ModuleType("pyspark.mllib.common")
    mllib.linalg = types.ModuleType("pyspark.mllib.linalg")
    ps.mllib = mllib
    for name, mod in (("pyspark", ps), ("pyspark.conf", conf), ("pyspark.serializers", ps.serializers), ("pyspark.rdd", ps.rdd),
                      ("pyspark.mllib", mllib), ("pyspark.mllib.common", mllib.common), ("pyspark.mllib.linalg", mllib.linalg)):
        sys.modules[name] = mod


def load_reference_twin():
    """The reference's module namespace with prepare_call_data / calculate_similarity_matrix / center_matrix defined."""
    src = REF.read_text()
    for pat, rep, want in REWRITES:
        src, cnt = re.subn(pat, rep, src)
        if cnt != want:
            raise SystemExit(f"the reference source changed: /{pat}/ matched {cnt} times, expected {want}")
    install_pyspark_stub()
    if not hasattr(numpy, "int"):
        numpy.int = int                       # what numpy.int was an alias of (the twin's :69)
    ns = {"__name__": "reference_variants_pca"}
    try:
        exec(compile(src, str(REF), "exec"), ns)          # the last statement, pca(sys.argv[1:]) (:201), needs the JVM
    except AttributeError as exc:
        if "no JVM behind this stand-in" not in str(exc):
            raise
    for f in ("prepare_call_data", "calculate_similarity_matrix", "center_matrix"):
        assert callable(ns.get(f)), f
    return ns


# ------------------------------------------------------------------------------------------------ cases
def make_variants(rng, ids, nv, allow_nocall=False, duplicate_every=0):
    """Variant records as the twin sees them (dicts with 'calls': [{'callSetId', 'genotype'}]); some without carriers, some
    without a 'calls' key at all (:33 `v.get('calls', [])`)."""
    out = []
    for j in range(nv):
        if j % 11 == 10:
            out.append({"referenceName": "17", "start": 41196311 + j})            # no 'calls'
            continue
        calls = []
        freq = rng.uniform(0.0, 0.6)
        for cid in ids:
            a, b = int(rng.random() < freq), int(rng.random() < freq)
            if rng.random() < 0.1:
                a *= 2                                                              # second alternate allele
            g = [a, b]
            if allow_nocall and rng.random() < 0.15:
                g = [-1, -1]
            calls.append({"callSetId": cid, "genotype": g})
        if duplicate_every and j % duplicate_every == 0:
            calls.append(dict(calls[j % len(calls)]))                               # a callset listed twice in one variant
        out.append({"referenceName": "17", "start": 41196311 + j, "calls": calls})
    return out


def run_case(ns, name, n, nv, partitions, seed, **kw):
    rng = numpy.random.default_rng(seed)
    ids = [f"set{seed}-{i}" for i in range(n)]
    id_to_index = {cid: i for i, cid in enumerate(ids)}
    variants = make_variants(rng, ids, nv, **kw)
    sc = SparkContext._active_spark_context
    py_rdd = sc.parallelize(variants, partitions)
    call_rdd = ns["prepare_call_data"](py_rdd, id_to_index)
    rows = call_rdd.collect()
    sim = ns["calculate_similarity_matrix"](call_rdd, n)
    sim_records = sorted(sim.collect())
    centered = ns["center_matrix"](sim, n).collect()
    assert len(sim_records) == n * n and len(centered) == n
    return {
        "name": name, "n": n, "partitions": partitions, "callset_ids": ids,
        "variants": variants,
        "call_rows": [[int(i) for i in r] for r in rows],
        "similarity_records": [[int(y), int(x), int(v)] for (y, x), v in sim_records],
        # doubles as hex strings: exact, whatever the JSON float printer does
        "centered_rows": [[[int(c), float(v).hex()] for c, v in row] for row in centered],
    }


def main():
    if not REF.exists():
        raise SystemExit(f"{REF} not found: the goldens can only be regenerated where the reference is checked out")
    ns = load_reference_twin()
    cases = [
        run_case(ns, "case_three_partitions", 12, 45, 3, seed=1),
        run_case(ns, "case_duplicates", 9, 30, 2, seed=2, duplicate_every=4),
        run_case(ns, "case_single_partition_wide", 31, 64, 1, seed=3),
        run_case(ns, "case_nocall_divergence", 7, 20, 2, seed=4, allow_nocall=True),
    ]
    OUT.write_text(json.dumps({"generated_from": str(REF), "rewrites": [[p, r, c] for p, r, c in REWRITES], "cases": cases},
                              separators=(",", ":")) + "\n")
    for c in cases:
        print(c["name"], "rows", len(c["call_rows"]), "S sum", sum(v for _, _, v in c["similarity_records"]))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
