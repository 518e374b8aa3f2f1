"""CsvDataSource (the host mirror of src/execution/datasource.rs:33-58) on the reference's fixtures. CPU only."""
import os

import numpy as np
import pytest

from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200 import host

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
CITIES = [("city", A.UTF8), ("lat", A.FLOAT64), ("lng", A.FLOAT64)]


def drain(ds):
    out = []
    while True:
        b = ds.next()
        if b is None:
            return out
        out.append(b)


def test_uk_cities_header_quirk_and_batching(golden):
    host.build()
    # has_headers = true unconditionally (datasource.rs:41): line 1 of the header-less file is dropped -> 36 rows
    b = drain(host.CsvDataSource(os.path.join(DATA, "uk_cities.csv"), CITIES, 1024))
    assert len(b) == 1 and len(b[0][0]) == 36
    c = golden["uk_cities"]
    assert b[0][0] == c["city"]  # quoted fields containing commas
    assert np.array_equal(b[0][1], np.array(c["lat"])) and np.array_equal(b[0][2], np.array(c["lng"]))
    # batch_size rows per next(), last batch ragged
    parts = drain(host.CsvDataSource(os.path.join(DATA, "uk_cities.csv"), CITIES, 10))
    assert [len(p[0]) for p in parts] == [10, 10, 10, 6]
    assert sum((p[0] for p in parts), []) == c["city"]


def test_typed_columns_and_errors(golden):
    b = drain(host.CsvDataSource(os.path.join(DATA, "aggregate_test_1.csv"), [("a", A.INT32), ("b", A.FLOAT64)], 1024))
    assert b[0][0].dtype == np.int32 and list(b[0][0]) == golden["aggregate_test_1"]["a"]
    assert list(b[0][1]) == golden["aggregate_test_1"]["b"]
    b = drain(host.CsvDataSource(os.path.join(DATA, "people.csv"), [("id", A.INT32), ("first_name", A.UTF8)], 1024))
    assert b[0][0][0] == 1 and b[0][1][0] == "Andy"
    with pytest.raises(host.ExecutionError):
        host.CsvDataSource(os.path.join(DATA, "does_not_exist.csv"), CITIES, 1024)
    with pytest.raises(host.ExecutionError) as e:  # a Utf8 field parsed as a number
        drain(host.CsvDataSource(os.path.join(DATA, "uk_cities.csv"), [("city", A.FLOAT64), ("lat", A.FLOAT64), ("lng", A.FLOAT64)], 1024))
    assert "ParseError" in e.value.msg
