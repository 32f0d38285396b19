"""CPU-side checks of the C++ host layer that need no GPU: Java Double.toString formatting."""
import math


def test_java_double_to_string_matches_oracle_and_known_values(native, oracle):
    from genestrip_b200 import host
    cases = {1.0: "1.0", 0.5: "0.5", 100.0: "100.0", 1234567.0: "1234567.0", 1.0e7: "1.0E7", 1.0e-3: "0.001", 1.0e-4: "1.0E-4",
             0.1: "0.1", 1 / 3: "0.3333333333333333", 123456789.125: "1.23456789125E8", 2.0e-5: "2.0E-5", 150.0: "150.0",
             -2.5: "-2.5", 1.7976931348623157e308: "1.7976931348623157E308"}
    for v, s in cases.items():
        assert host.java_double_to_string(v) == s
    import random
    rng = random.Random(1)
    for _ in range(2000):
        v = rng.random() * 10 ** rng.randint(-8, 12)
        assert host.java_double_to_string(v) == oracle.java_double_to_string(v)
        assert float(host.java_double_to_string(v).replace("E", "e")) == v
    assert host.java_double_to_string(math.nan) == "NaN" and host.java_double_to_string(math.inf) == "Infinity"
