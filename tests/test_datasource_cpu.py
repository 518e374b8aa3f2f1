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


def test_empty_fields_become_nulls_and_boolean_columns():
    # test/data/null_test.csv of the reference (no reference test reads it: semantics restated from the
    # arrow 0.12 csv reader — empty primitive field -> null, empty Utf8 field -> "", bool via str::parse)
    fields = [("c_int", A.INT32), ("c_float", A.FLOAT64), ("c_string", A.UTF8), ("c_bool", A.BOOL)]
    # has_headers = true: the header line is the one that is dropped here
    (b,) = drain(host.CsvDataSource(os.path.join(DATA, "null_test.csv"), fields, 1024))
    assert list(b[0]) == [1, 2, 3, 4, 5]
    vals, mask = b[1]
    assert list(mask) == [True, True, False, True, True]
    assert list(vals[mask]) == [1.1, 2.2, 4.4, 6.6] and vals[2] == 0.0
    assert b[2] == ["1.11", "2.22", "3.33", "", ""]
    assert b[3].dtype == bool and list(b[3]) == [True, True, True, False, False]
    # ragged batches keep validity aligned
    parts = drain(host.CsvDataSource(os.path.join(DATA, "null_test.csv"), fields, 2))
    assert [len(p[0]) for p in parts] == [2, 2, 1]
    assert isinstance(parts[1][1], tuple) and list(parts[1][1][1]) == [False, True]
    assert not isinstance(parts[0][1], tuple)  # no nulls in the first batch: no bitmap


@pytest.mark.parametrize("text,dtype", [("1,abc\n", A.INT32), ("1,300\n", A.INT8), ("1,-1\n", A.UINT16), ("1, 5\n", A.INT64),
                                        ("1,5x\n", A.FLOAT64), ("1,yes\n", A.BOOL), ("1,99999999999999999999\n", A.INT64)])
def test_unparsable_values_are_parse_errors(tmp_path, text, dtype):
    f = tmp_path / "t.csv"
    f.write_text("a,b\n" + text)
    with pytest.raises(host.ExecutionError) as e:
        drain(host.CsvDataSource(str(f), [("a", A.INT32), ("b", dtype)], 16))
    assert "ParseError" in e.value.msg and "line 2" in e.value.msg
