"""The oracle and the host mirror against vectors produced by THE REFERENCE'S OWN CODE (tests/golden/reference_twin.json,
made by tests/golden/make_reference_twin_golden.py from /root/reference/src/main/python/variants_pca.py: the reference's
Python twin of the hot path, its `prepare_call_data` :19-52, `calculate_similarity_matrix` :54-82 and `center_matrix` :84-121
executed on an in-memory RDD stand-in).  These are the only reference-produced numbers that exist for this path: the encode,
similarity and centering steps are pinned by them bit for bit; the eigen step (`perform_pca`, :123-152) needs the JVM and stays
pinned on upstream's RowMatrixSuite known-answer case (tests/test_oracle.py)."""
import json
from pathlib import Path

import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200.variants_pca import VariantsPcaDriver

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "reference_twin.json").read_text())
CASES = {c["name"]: c for c in GOLD["cases"]}
SCALA_RULE_CASES = [n for n in CASES if n != "case_nocall_divergence"]     # no negative allele: `any(g)` == `_ > 0`


def _dense(case):
    n = case["n"]
    S = np.zeros((n, n), np.int64)
    for y, x, v in case["similarity_records"]:
        S[y, x] = v
    return S


def _centered(case):
    n = case["n"]
    C = np.full((n, n), np.nan)
    for row, cols in enumerate(case["centered_rows"]):
        for col, hexval in cols:
            C[row, col] = float.fromhex(hexval)
    assert not np.isnan(C).any()
    return C


def _calls(case):
    return [None if "calls" not in v else [(c["callSetId"], c["genotype"]) for c in v["calls"]] for v in case["variants"]]


def _csr(rows):
    off = np.zeros(len(rows) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in rows])
    return off, np.asarray([i for r in rows for i in r], np.int32)


@pytest.mark.parametrize("name", SCALA_RULE_CASES)
def test_encode_equals_the_reference_twin(oracle, name):
    """variants_pca.py:33-50 == VariantsPca.scala:56-60 + :163-167 when no allele is negative: the call rows, in order."""
    case = CASES[name]
    mapping = {cid: i for i, cid in enumerate(case["callset_ids"])}
    assert oracle.np_get_calls(_calls(case), mapping) == case["call_rows"]
    # the host mirror of the driver (records -> getCallsRdd), same rows
    variants = [pkg.Variant("17", start=v["start"], end=v["start"] + 1, referenceBases="A", alternateBases=["G"],
                            calls=None if "calls" not in v else [pkg.Call(c["callSetId"], genotype=c["genotype"]) for c in v["calls"]])
                for v in case["variants"]]
    conf = pkg.PcaConf([])
    callsets = [(cid, cid.upper()) for cid in case["callset_ids"]]
    d = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=callsets, datasets=[variants]))
    assert d.getCallsRdd(d.getData).collect() == case["call_rows"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_similarity_equals_the_reference_twin(oracle, name):
    """variants_pca.py:68-82 (`matrix[y][x] += 1` per partition, reduceByKey) == VariantsPca.scala:182-191: all N^2 counts."""
    case = CASES[name]
    n, rows = case["n"], case["call_rows"]
    S = _dense(case)
    assert np.array_equal(S, S.T)
    assert np.array_equal(oracle.np_similarity(n, rows), S)
    off, idx = _csr(rows)
    for parts in (1, case["partitions"], 5):
        assert np.array_equal(oracle.c_similarity(n, off, idx, parts), S)
    assert np.array_equal(oracle.c_similarity_stream(n, off, idx), S)


@pytest.mark.parametrize("name", sorted(CASES))
def test_centering_equals_the_reference_twin_bit_for_bit(oracle, name):
    """variants_pca.py:99-121 (row sums, `float(sum) / N / N`, `val - row_mean - col_mean + matrix_mean`) ==
    VariantsPca.scala:199-223: every centred entry, bit for bit."""
    case = CASES[name]
    S = _dense(case).astype(np.int32)
    want = _centered(case)
    C_np, row_sums, nz = oracle.np_center(S)
    C_c, _, _ = oracle.c_center(S)
    assert np.array_equal(C_np, want) and np.array_equal(C_c, want)
    assert np.array_equal(row_sums, S.sum(axis=1))


def test_the_twins_any_rule_differs_from_scala_on_no_calls(oracle):
    """The one divergence inside the reference: the twin keeps a call when `any(genotype)` (:36), so a no-call (-1, -1) counts
    as variation; Scala's `_ > 0` (VariantsPca.scala:58) does not -- this repository follows Scala.  Pin what each does."""
    case = CASES["case_nocall_divergence"]
    mapping = {cid: i for i, cid in enumerate(case["callset_ids"])}
    calls = _calls(case)
    scala_rows = oracle.np_get_calls(calls, mapping)
    twin_rule = [[mapping[cid] for cid, g in (c or []) if any(g)] for c in calls]
    twin_rule = [r for r in twin_rule if r]
    assert twin_rule == case["call_rows"] and twin_rule != scala_rows
    nocall = {(j, mapping[cid]) for j, c in enumerate(calls) for cid, g in (c or []) if g == [-1, -1]}
    assert nocall and sum(len(r) for r in twin_rule) - sum(len(r) for r in scala_rows) == len(nocall)
