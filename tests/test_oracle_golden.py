"""Pin the CPU oracle (oracle/gfa_oracle.c + SciPy) against golden vectors recorded from the
real reference by tools/gen_golden.py.  CPU only."""
import base64

import pytest

import golden_inputs as gi
import parity_util as pu
from oracle.oracle import oracle_convert_format, oracle_parse_gfa

CASES = pu.load_json("cases.json")
FUZZ = pu.load_json("fuzz.json")
DRB1 = pu.load_json("drb1.json")


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_literal_cases(case):
    text = base64.b64decode(case["text_b64"])
    assert text == dict(gi.LITERAL_CASES)[case["name"]], "golden file is stale: rerun tools/gen_golden.py"
    for run in case["runs"]:
        pu.run_and_compare(oracle_parse_gfa, oracle_convert_format, text, run["mode"], run["expect"],
                           full=True, what=f"{case['name']} {run['mode']}")


@pytest.mark.parametrize("entry", FUZZ, ids=[str(e["seed"]) for e in FUZZ])
def test_oracle_fuzz(entry):
    text = gi.fuzz_text(entry["seed"])
    assert len(text) == entry["nbytes"]
    for run in entry["runs"]:
        pu.run_and_compare(oracle_parse_gfa, oracle_convert_format, text, run["mode"], run["expect"],
                           full=False, what=f"fuzz{entry['seed']} {run['mode']}")


def test_oracle_drb1_fixture():
    path = pu.GOLD / "DRB1-3123_unsorted.gfa"
    for run in DRB1["runs"]:
        pu.run_and_compare(oracle_parse_gfa, oracle_convert_format, path, run["mode"], run["expect"],
                           full=False, what=f"drb1 {run['mode']}")
    # tripwire from SURVEY.md 8(c): default directed CSR fingerprint
    assert DRB1["runs"][0]["expect"]["csr"]["sha"] == "a7f5b4532a8d7075"
