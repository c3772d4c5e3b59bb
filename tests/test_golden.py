"""The oracle against the committed golden vectors (tests/golden/golden.json), which were produced by
the reference library itself. Runs anywhere; does not need /root/reference."""
import hashlib

import numpy as np

from helpers import golden_cases, golden_check_stream, golden_image, golden_kwargs, oracle

SURVEY_KATS = {  # SURVEY.md section 8a, hex of the complete qb3_encode output
    "K1": "5142338007000700000008534308002376fbaed98c54014454436a9b4933e34452db4c9a192792da66d2cc3891d4369366c609",
    "K4": "5142338007000700000008534308002376fbaed98c5401445400",
    "K5": "514233800700070002000843420300010101534308002376fbaed98c54014454ff01000000000000803f11091109118988d09f"
          "00000000000000e07bd5aa55dbd6472e3189495403ad56addab600",
    "K6": "5142338007000700000205534308002376fbaed98c54014454372b597773d9dc79b92d597773d9dc79b9854c65b298ac52f12ec9ba9bcbe6ce0b",
    "K7": "51423380080006000005045156010003534308002376fbaed98c5401445413611008411882412882c0084a9f4993f1325a27b0"
          "9864c9b41472a28620428c304409c184aa3369325e46eb0416932c99964206",
    "K8": "514233803f003f00000007534308002376fbaed98c54014454ffff3c",
}


def test_golden_file_matches_survey_kats():
    by_name = {c["name"]: c for c in golden_cases()}
    for k, hx in SURVEY_KATS.items():
        assert by_name[k]["stream"] == hx, k


def test_oracle_encode_matches_golden():
    O = oracle()
    for case in golden_cases():
        golden_check_stream(case, O.encode(golden_image(case), **golden_kwargs(case)))


def test_oracle_decode_matches_golden():
    O = oracle()
    for case in golden_cases():
        if case["kind"] != "small":
            continue
        d = O.decode(bytes.fromhex(case["stream"]), identity_default=False)  # as the reference decoder does
        if case["ref_decoded"] is None:
            assert d is None, case["name"]
        else:
            assert hashlib.sha256(d.tobytes()).hexdigest() == case["ref_decoded"], case["name"]
        if case.get("quanta", 1) == 1 and d is not None:
            assert np.array_equal(O.decode(bytes.fromhex(case["stream"])), golden_image(case)), case["name"]
