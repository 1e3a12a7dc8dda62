"""Shared test workloads: the reference's own test tables / queries, restated, plus extra cases.

Tables and literal expected rows come from the reference's tests (data only):
  FRUITS            tests/test_execution.py:17-27
  USERS / ORDERS    tests/test_e2e.py:22-56 (schemas :59-79)
  SQL_CASES         tests/test_e2e.py:88-419  (20 SQL -> rows golden vectors)
  DF_CASES          tests/test_execution.py:33-288 (DataFrame API golden vectors)
Query builders take a namespace ``ns`` exposing DataFrame / Col / Lit / F so the same definition can
be built with this package's classes or (in the build container) with the real reference classes.
"""

from __future__ import annotations

import types
from datetime import datetime
from pathlib import Path
from typing import Any, Callable

to_date = datetime.fromisoformat

FRUITS = [
    {"fruit": "apple", "quantity": 3, "color": "red"},
    {"fruit": "banana", "quantity": 5, "color": "yellow"},
    {"fruit": "orange", "quantity": 2, "color": "orange"},
    {"fruit": "apple", "quantity": 4, "color": "green"},
    {"fruit": "banana", "quantity": 7, "color": "yellow"},
]
FRUITS_PRICED = [  # examples/fruit_aggregation.py / README.md:97-106 (BASELINE config 1)
    {"fruit": "apple", "quantity": 3, "color": "red", "price": 1.5},
    {"fruit": "banana", "quantity": 5, "color": "yellow", "price": 1.9},
    {"fruit": "orange", "quantity": 2, "color": "orange", "price": 1.2},
    {"fruit": "orange", "quantity": 4, "color": "orange", "price": 2.2},
]

USERS_SCHEMA = ["user_id:INT", "first_name:STR", "last_name:STR", "age:INT", "country:STR"]
USERS = [
    (1, "Alice", "Smith", 25, "USA"), (2, "Bob", "Johnson", 30, "Canada"), (3, "Charlie", "Brown", 22, "USA"),
    (4, "David", "Wilson", 35, "UK"), (5, "Eva", "Davis", 28, "Canada"), (6, "Frank", "Miller", 40, "USA"),
    (7, "Grace", "Taylor", 27, "UK"), (8, "Hank", "Anderson", 32, "USA"), (9, "Ivy", "Thomas", 26, "Canada"),
    (10, "Jack", "Jackson", 24, "USA"), (11, "Kate", "White", 29, "UK"), (12, "Leo", "Harris", 33, "USA"),
    (13, "Mia", "Martin", 31, "Canada"), (14, "Nick", "Thompson", 23, "UK"), (15, "Olivia", "Garcia", 36, "USA"),
]
ORDERS_SCHEMA = ["order_id:INT", "user_id:INT", "product:STR", "quantity:INT", "price:FLOAT", "order_date:TIMESTAMP"]
ORDERS = [
    (1, 1, "Laptop", 1, 1200.0, "2025-01-01"), (2, 2, "Mouse", 2, 25.0, "2025-01-05"),
    (3, 3, "Keyboard", 1, 45.0, "2025-02-10"), (4, 1, "Monitor", 2, 300.0, "2025-03-15"),
    (5, 4, "Laptop", 1, 1100.0, "2025-03-20"), (6, 5, "Mouse", 1, 30.0, "2025-04-01"),
    (7, 6, "Keyboard", 2, 50.0, "2025-04-10"), (8, 7, "Monitor", 1, 280.0, "2025-05-05"),
    (9, 8, "Laptop", 1, 1300.0, "2025-05-10"), (10, 9, "Mouse", 3, 27.0, "2025-06-01"),
    (11, 10, "Keyboard", 1, 40.0, "2025-06-15"), (12, 11, "Monitor", 2, 290.0, "2025-07-01"),
    (13, 12, "Laptop", 1, 1250.0, "2025-07-10"), (14, 13, "Mouse", 2, 26.0, "2025-07-15"),
    (15, 14, "Keyboard", 1, 42.0, "2025-08-01"),
]


def namespace(which: str = "mirror") -> types.SimpleNamespace:
    """Classes to build queries with: this package's mirror, or the real reference (container only)."""
    if which == "mirror":
        from minispark_b200.constants import ColumnType
        from minispark_b200.dataframe import DataFrame
        from minispark_b200.io import BlockFile
        from minispark_b200.sql import Col, Functions, Lit
    else:
        from mini_spark.constants import ColumnType  # type: ignore[import-not-found]
        from mini_spark.dataframe import DataFrame  # type: ignore[import-not-found]
        from mini_spark.io import BlockFile  # type: ignore[import-not-found]
        from mini_spark.sql import Col, Functions, Lit  # type: ignore[import-not-found]
    return types.SimpleNamespace(DataFrame=DataFrame, Col=Col, Lit=Lit, F=Functions, BlockFile=BlockFile, ColumnType=ColumnType)


def _schema(ns: Any, spec: list[str]) -> list:
    names = {"INT": ns.ColumnType.INTEGER, "STR": ns.ColumnType.STRING, "FLOAT": ns.ColumnType.FLOAT,
             "TIMESTAMP": ns.ColumnType.TIMESTAMP}
    return [(s.split(":")[0], names[s.split(":")[1]]) for s in spec]


def write_tables(folder: Path, ns: Any = None) -> dict[str, str]:
    """Write fruits / fruits_priced / users / orders BlockFiles; returns name -> path."""
    ns = ns or namespace()
    folder.mkdir(parents=True, exist_ok=True)
    ns.BlockFile(folder / "fruits.bin").write_rows([dict(r) for r in FRUITS])
    ns.BlockFile(folder / "fruits_priced.bin").write_rows([dict(r) for r in FRUITS_PRICED])
    ns.BlockFile(folder / "users", _schema(ns, USERS_SCHEMA)).write_data(tuple(map(list, zip(*USERS))))
    ns.BlockFile(folder / "orders", _schema(ns, ORDERS_SCHEMA)).write_data(tuple(map(list, zip(*ORDERS))))
    return {"fruits": str(folder / "fruits.bin"), "fruits_priced": str(folder / "fruits_priced.bin"),
            "users": str(folder / "users"), "orders": str(folder / "orders")}


def rows(schema: tuple[str, ...], data: list[tuple]) -> list[dict]:
    return [dict(zip(schema, r)) for r in data]


# ---------------------------------------------------------------------------------------------------
# SQL -> rows golden vectors of the reference (tests/test_e2e.py:88-419)
# ---------------------------------------------------------------------------------------------------
_ORDER_COLS = ("order_id", "user_id", "product", "quantity", "price", "order_date")


def _orders(ids: list[int]) -> list[tuple]:
    return [(o[0], o[1], o[2], o[3], o[4], to_date(o[5])) for o in ORDERS if o[0] in ids]


SQL_CASES: list[tuple[str, str, list[dict]]] = [
    ("select_star", "SELECT * FROM '{users}';", rows(("user_id", "first_name", "last_name", "age", "country"), USERS)),
    ("where_string_eq", "SELECT first_name, last_name FROM '{users}' WHERE country='USA';",
     rows(("first_name", "last_name"), [(u[1], u[2]) for u in USERS if u[4] == "USA"])),
    ("concat", "SELECT first_name + ' ' + last_name AS full_name FROM '{users}';",
     rows(("full_name",), [(f"{u[1]} {u[2]}",) for u in USERS])),
    ("int_arith", "SELECT user_id, age, age+5 AS age_in_5_years FROM '{users}';",
     rows(("user_id", "age", "age_in_5_years"), [(u[0], u[3], u[3] + 5) for u in USERS])),
    ("float_gt", "SELECT * FROM '{orders}' WHERE price > 100;", rows(_ORDER_COLS, _orders([1, 4, 5, 8, 9, 12, 13]))),
    ("int_times_float", "SELECT product, quantity*price AS total_value FROM '{orders}';",
     rows(("product", "total_value"), [(o[2], o[3] * o[4]) for o in ORDERS])),
    ("between_ts", "SELECT * FROM '{orders}' WHERE order_date BETWEEN '2025-03-01' AND '2025-06-01';",
     rows(_ORDER_COLS, _orders([4, 5, 6, 7, 8, 9, 10]))),
    ("like", "SELECT * FROM '{orders}' WHERE product LIKE '%top%';", rows(_ORDER_COLS, _orders([1, 5, 9, 13]))),
    ("group_count", "SELECT country, COUNT() AS user_count FROM '{users}' GROUP BY country;",
     rows(("country", "user_count"), [("USA", 7), ("Canada", 4), ("UK", 4)])),
    ("group_sum_expr", "SELECT user_id, SUM(quantity*price) AS total_spent FROM '{orders}' GROUP BY user_id;",
     rows(("user_id", "total_spent"), [(1, 1800.0), (2, 50.0), (3, 45.0), (4, 1100.0), (5, 30.0), (6, 100.0), (7, 280.0),
                                      (8, 1300.0), (9, 81.0), (10, 40.0), (11, 580.0), (12, 1250.0), (13, 52.0), (14, 42.0)])),
    ("group_avg_float", "SELECT product, AVG(price) AS avg_price FROM '{orders}' GROUP BY product;",
     rows(("product", "avg_price"), [("Laptop", (1200 + 1100 + 1300 + 1250) / 4), ("Mouse", (25 + 30 + 27 + 26) / 4),
                                     ("Keyboard", (45 + 50 + 40 + 42) / 4), ("Monitor", (300 + 280 + 290) / 3)])),
    ("group_avg_int", "SELECT country, AVG(age) AS avg_age FROM '{users}' GROUP BY country;",
     rows(("country", "avg_age"), [("USA", (25 + 22 + 40 + 32 + 24 + 33 + 36) / 7), ("Canada", (30 + 28 + 26 + 31) / 4),
                                   ("UK", (35 + 27 + 29 + 23) / 4)])),
    ("having", "SELECT user_id, COUNT() AS order_count FROM '{orders}' GROUP BY user_id HAVING COUNT() > 1;",
     rows(("user_id", "order_count"), [(1, 2)])),
    ("join", "SELECT u.first_name, o.product FROM '{users}' AS u JOIN '{orders}' AS o ON u.user_id=o.user_id;",
     rows(("first_name", "product"), [(next(u[1] for u in USERS if u[0] == o[1]), o[2]) for o in ORDERS])),
    ("join_group_count", "SELECT u.country, COUNT() AS orders_count FROM '{users}' AS u JOIN '{orders}' AS o "
     "ON u.user_id=o.user_id GROUP BY u.country;", rows(("country", "orders_count"), [("USA", 7), ("Canada", 4), ("UK", 4)])),
    ("join_group_sum", "SELECT u.first_name, SUM(o.quantity*o.price) AS spent FROM '{users}' AS u JOIN '{orders}' AS o "
     "ON u.user_id=o.user_id GROUP BY u.first_name;",
     rows(("first_name", "spent"), [("Alice", 1800.0), ("Bob", 50.0), ("Charlie", 45.0), ("David", 1100.0), ("Eva", 30.0),
                                    ("Frank", 100.0), ("Grace", 280.0), ("Hank", 1300.0), ("Ivy", 81.0), ("Jack", 40.0),
                                    ("Kate", 580.0), ("Leo", 1250.0), ("Mia", 52.0), ("Nick", 42.0)])),
    ("left_join_where", "SELECT u.first_name, o.product, o.price FROM '{users}' AS u LEFT JOIN '{orders}' AS o "
     "ON u.user_id=o.user_id WHERE o.price > 100;",
     rows(("first_name", "product", "price"), [("Alice", "Laptop", 1200.0), ("Alice", "Monitor", 300.0), ("David", "Laptop", 1100.0),
                                               ("Grace", "Monitor", 280.0), ("Hank", "Laptop", 1300.0), ("Kate", "Monitor", 290.0),
                                               ("Leo", "Laptop", 1250.0)])),
    ("left_join_where_ts", "SELECT u.first_name, o.product, o.order_date FROM '{orders}' AS o LEFT JOIN '{users}' AS u "
     "ON u.user_id=o.user_id WHERE o.order_date > '2025-05-01';",
     rows(("first_name", "product", "order_date"),
          [("Hank", "Laptop", to_date("2025-05-10")), ("Grace", "Monitor", to_date("2025-05-05")), ("Ivy", "Mouse", to_date("2025-06-01")),
           ("Jack", "Keyboard", to_date("2025-06-15")), ("Kate", "Monitor", to_date("2025-07-01")), ("Leo", "Laptop", to_date("2025-07-10")),
           ("Mia", "Mouse", to_date("2025-07-15")), ("Nick", "Keyboard", to_date("2025-08-01"))])),
    ("sum_and_max", "SELECT product, SUM(quantity) AS total_quantity, MAX(price) AS max_price FROM '{orders}' GROUP BY product;",
     rows(("product", "total_quantity", "max_price"), [("Laptop", 4, 1300.0), ("Mouse", 8, 30.0), ("Keyboard", 5, 50.0), ("Monitor", 5, 300.0)])),
    ("join_having", "SELECT u.country, COUNT() AS orders_count, SUM(o.quantity*o.price) AS total_sales FROM '{users}' AS u "
     "JOIN '{orders}' AS o ON u.user_id=o.user_id GROUP BY u.country HAVING SUM(o.quantity*o.price) > 500;",
     rows(("country", "orders_count", "total_sales"), [("USA", 7, 4535.0), ("UK", 4, 2002.0)])),
]
# scan / filter / project queries must keep input order (tests/test_execution.py:40-46,139-143)
ORDERED_SQL = {"select_star", "where_string_eq", "concat", "int_arith", "float_gt", "int_times_float", "between_ts", "like"}


# ---------------------------------------------------------------------------------------------------
# DataFrame-API cases (tests/test_execution.py:33-288 + README fruit example + extra coverage)
# name -> (builder(ns, tables, engine) -> DataFrame, expected rows or None, ordered?)
# ---------------------------------------------------------------------------------------------------
def _df(ns: Any, engine: Any, table: str) -> Any:
    return ns.DataFrame(engine).table(table)


DF_CASES: dict[str, tuple[Callable[[Any, dict, Any], Any], Any, bool]] = {
    "table_load": (lambda ns, t, e: _df(ns, e, t["fruits"]), [dict(r) for r in FRUITS], True),
    "select": (lambda ns, t, e: _df(ns, e, t["fruits"]).select(ns.Col("fruit")), [{"fruit": r["fruit"]} for r in FRUITS], True),
    "select_expression": (lambda ns, t, e: _df(ns, e, t["fruits"]).select(ns.Col("quantity") + 3),
                          [{"quantity_add_lit_3": r["quantity"] + 3} for r in FRUITS], True),
    "select_alias": (lambda ns, t, e: _df(ns, e, t["fruits"]).select(ns.Col("fruit").alias("fruit_name")),
                     [{"fruit_name": r["fruit"]} for r in FRUITS], True),
    "select_star": (lambda ns, t, e: _df(ns, e, t["fruits"]).select(ns.Col("*")), [dict(r) for r in FRUITS], True),
    "filter": (lambda ns, t, e: _df(ns, e, t["fruits"]).filter(ns.Col("quantity") > 3),
               [dict(r) for r in FRUITS if r["quantity"] > 3], True),
    "groupby_count": (lambda ns, t, e: _df(ns, e, t["fruits"]).group_by(ns.Col("fruit")).agg(ns.F.count()),
                      [{"fruit": "apple", "count": 2}, {"fruit": "banana", "count": 2}, {"fruit": "orange", "count": 1}], False),
    "groupby_multi": (lambda ns, t, e: _df(ns, e, t["fruits"]).group_by(ns.Col("fruit")).agg(
        ns.F.count(), ns.F.min(ns.Col("quantity")).alias("min"), ns.F.max(ns.Col("quantity")).alias("max"),
        ns.F.sum(ns.Col("quantity")).alias("sum")),
        [{"fruit": "apple", "count": 2, "min": 3, "max": 4, "sum": 7}, {"fruit": "banana", "count": 2, "min": 5, "max": 7, "sum": 12},
         {"fruit": "orange", "count": 1, "min": 2, "max": 2, "sum": 2}], False),
    "self_join": (lambda ns, t, e: _df(ns, e, t["fruits"]).select(ns.Col("fruit").alias("fruit_left"), ns.Col("color")).join(
        ns.DataFrame().table(t["fruits"]).select(ns.Col("fruit").alias("fruit_right"), ns.Col("quantity")),
        on=ns.Col("fruit_left") == ns.Col("fruit_right"), how="inner"),
        [{"fruit_left": a["fruit"], "color": a["color"], "fruit_right": b["fruit"], "quantity": b["quantity"]}
         for b in FRUITS for a in FRUITS if a["fruit"] == b["fruit"]], False),
    # README.md:97-106 / examples/fruit_aggregation.py: apple 4.5, banana 9.5, orange 11.2 (f32)
    "fruit_aggregation": (lambda ns, t, e: _df(ns, e, t["fruits_priced"]).group_by(ns.Col("fruit")).agg(
        ns.F.sum(ns.Col("quantity") * ns.Col("price")).alias("total_price")),
        # exact f32 values as the reference returns them (11.2 in the README is display rounding of 11.200000762939453)
        [{"fruit": "apple", "total_price": 4.5}, {"fruit": "banana", "total_price": 9.5}, {"fruit": "orange", "total_price": 11.200000762939453}], False),
    # extra coverage (expected rows come from the oracle / the real reference, see tests/golden)
    "int_ops": (lambda ns, t, e: _df(ns, e, t["users"]).select(
        ns.Col("user_id"), (ns.Col("age") // 7).alias("fd"), (ns.Col("age") % 7).alias("md"),
        ((ns.Col("age") - 30) // 4).alias("nfd"), ((ns.Col("age") - 30) % 4).alias("nmd"), (ns.Col("age") / 4).alias("td"),
        (ns.Col("age") * ns.Col("user_id") - 3).alias("mul")), None, True),
    "float_ops": (lambda ns, t, e: _df(ns, e, t["orders"]).select(
        ns.Col("order_id"), (ns.Col("price") // 7).alias("fd"), (ns.Col("price") % 7).alias("md"),
        ((ns.Col("price") - 300) // 7).alias("nfd"), ((ns.Col("price") - 300) % 7).alias("nmd"),
        (ns.Col("price") / ns.Col("quantity")).alias("td")), None, True),
    "filter_or_and": (lambda ns, t, e: _df(ns, e, t["users"]).filter(
        ((ns.Col("age") > 30) & (ns.Col("age") <= 36)) | (ns.Col("user_id") == 1)).select(ns.Col("first_name"), ns.Col("age")), None, True),
    "filter_ne_string": (lambda ns, t, e: _df(ns, e, t["users"]).filter(ns.Col("country") != "USA").select(ns.Col("first_name")), None, True),
    "filter_missing_string": (lambda ns, t, e: _df(ns, e, t["users"]).filter(ns.Col("country") == "Mars"), None, True),
    "like_underscore": (lambda ns, t, e: _df(ns, e, t["users"]).filter(ns.Col("first_name").like("_a%")).select(ns.Col("first_name")), None, True),
    "min_max_float": (lambda ns, t, e: _df(ns, e, t["orders"]).group_by(ns.Col("product")).agg(
        ns.F.min(ns.Col("price")).alias("lo"), ns.F.max(ns.Col("price")).alias("hi"), ns.F.avg(ns.Col("quantity")).alias("aq")), None, False),
    "group_by_int_expr": (lambda ns, t, e: _df(ns, e, t["users"]).group_by((ns.Col("age") // 10).alias("decade")).agg(
        ns.F.count(), ns.F.sum(ns.Col("age")).alias("ages")), None, False),
    "group_by_float": (lambda ns, t, e: _df(ns, e, t["orders"]).group_by(ns.Col("price")).agg(ns.F.count()), None, False),
    "group_by_timestamp": (lambda ns, t, e: _df(ns, e, t["orders"]).group_by(ns.Col("order_date")).agg(ns.F.sum(ns.Col("quantity")).alias("q")), None, False),
    "filter_then_group": (lambda ns, t, e: _df(ns, e, t["orders"]).filter(ns.Col("quantity") > 1).group_by(ns.Col("product")).agg(
        ns.F.sum(ns.Col("price") * ns.Col("quantity")).alias("v"), ns.F.count()), None, False),
    "project_then_filter": (lambda ns, t, e: _df(ns, e, t["orders"]).select(
        ns.Col("product"), (ns.Col("price") * ns.Col("quantity")).alias("v")).filter(ns.Col("v") > 100), None, True),
    "join_string_keys": (lambda ns, t, e: _df(ns, e, t["fruits"]).select(ns.Col("color").alias("c1"), ns.Col("quantity").alias("q1")).join(
        ns.DataFrame().table(t["fruits_priced"]).select(ns.Col("color").alias("c2"), ns.Col("price")),
        on=ns.Col("c1") == ns.Col("c2"), how="inner"), None, False),
    "join_then_filter_group": (lambda ns, t, e: ns.DataFrame(e).table(t["users"]).alias("u").join(
        ns.DataFrame().table(t["orders"]).alias("o"), on=ns.Col("u.user_id") == ns.Col("o.user_id"), how="inner")
        .filter(ns.Col("o.quantity") >= 1).group_by(ns.Col("u.country")).agg(
            ns.F.sum(ns.Col("o.price")).alias("sales"), ns.F.avg(ns.Col("u.age")).alias("avg_age"), ns.F.count()), None, False),
    # arithmetic shapes of the register-resident interpreter (operand order of -, fused column / constant forms, CSE)
    "agg_arith_shapes": (lambda ns, t, e: _df(ns, e, t["orders"]).filter(ns.Col("quantity") >= 1).group_by(ns.Col("product")).agg(
        ns.F.sum(ns.Col("price") - ns.Col("quantity")).alias("a"), ns.F.sum(ns.Col("quantity") - ns.Col("price")).alias("b"),
        ns.F.sum(ns.Lit(100) - ns.Col("price")).alias("c"), ns.F.sum(ns.Col("price") - 100).alias("d"),
        ns.F.sum((ns.Col("price") + ns.Col("quantity")) * (ns.Col("price") - ns.Col("quantity"))).alias("e"),
        ns.F.sum(ns.Col("price") * ns.Col("price") * ns.Col("quantity")).alias("f"),
        ns.F.min(ns.Col("price") - 1).alias("g"), ns.F.max(ns.Lit(2) * ns.Col("price")).alias("h"),
        ns.F.sum((ns.Col("price") + ns.Col("quantity")) * (ns.Col("price") - ns.Col("quantity")) * ns.Col("price")).alias("i"),
        ns.F.avg(ns.Col("price") * (ns.Lit(1) - ns.Col("quantity")) * (ns.Lit(1) + ns.Col("price"))).alias("j"),
        ns.F.sum(ns.Col("quantity")).alias("k"), ns.F.max(ns.Col("quantity")).alias("l"), ns.F.count()), None, False),
    "agg_float_filter": (lambda ns, t, e: _df(ns, e, t["orders"]).filter(ns.Col("price") <= 300.0).filter(ns.Col("order_date") > "2025-01-01")
        .group_by(ns.Col("product")).agg(ns.F.sum(ns.Col("price")).alias("s"), ns.F.min(ns.Col("price")).alias("lo")), None, False),
    # STRING ordering (Python string order, sql.py:262-266): against a literal, literal on the left, between two columns
    "string_order_literal": (lambda ns, t, e: _df(ns, e, t["users"]).filter((ns.Col("first_name") >= "F") & (ns.Col("last_name") < "Smith"))
                             .select(ns.Col("first_name"), ns.Col("last_name")), None, True),
    "string_order_columns": (lambda ns, t, e: _df(ns, e, t["users"]).filter(ns.Col("first_name") > ns.Col("last_name"))
                             .filter(ns.Col("country") <= ns.Col("first_name")).select(ns.Col("first_name"), ns.Col("last_name"), ns.Col("country")), None, True),
    "string_order_group": (lambda ns, t, e: _df(ns, e, t["orders"]).filter(ns.Lit("Laptop") < ns.Col("product")).group_by(ns.Col("product"))
                           .agg(ns.F.count(), ns.F.sum(ns.Col("quantity")).alias("q")), None, False),
    "empty_result": (lambda ns, t, e: _df(ns, e, t["orders"]).filter(ns.Col("price") > 1e9), [], True),
    "concat_filter": (lambda ns, t, e: _df(ns, e, t["users"]).filter(ns.Col("age") < 30).select(
        (ns.Col("first_name") + "-" + ns.Col("country")).alias("tag"), ns.Col("age")), None, True),
}


def q1(ns: Any, table: str, engine: Any = None) -> Any:
    """TPC-H Q1 as the reference benchmarks it (examples/benchmark.py:51-68), via the DataFrame API."""
    C, F, L = ns.Col, ns.F, ns.Lit
    return (ns.DataFrame(engine).table(table).filter(C("l_shipdate") <= "1998-12-01").group_by(C("l_returnflag")).agg(
        F.sum(C("l_quantity")).alias("sum_qty"),
        F.sum(C("l_extendedprice")).alias("sum_base_price"),
        F.sum(C("l_extendedprice") * (L(1) - C("l_discount"))).alias("sum_disc_price"),
        F.sum(C("l_extendedprice") * (L(1) - C("l_discount")) * (L(1) + C("l_tax"))).alias("sum_charge"),
        F.avg(C("l_quantity")).alias("avg_qty"),
        F.avg(C("l_extendedprice")).alias("avg_price"),
        F.avg(C("l_discount")).alias("avg_disc"),
        F.count().alias("count_order")))


Q1_SQL = """
SELECT
    l_returnflag,
    SUM(l_quantity)        AS sum_qty,
    SUM(l_extendedprice)   AS sum_base_price,
    SUM(l_extendedprice * (1 - l_discount))              AS sum_disc_price,
    SUM(l_extendedprice * (1 - l_discount) * (1 + l_tax)) AS sum_charge,
    AVG(l_quantity)        AS avg_qty,
    AVG(l_extendedprice)   AS avg_price,
    AVG(l_discount)        AS avg_disc,
    COUNT()               AS count_order
FROM
     '{table}'
WHERE
    l_shipdate <= '1998-12-01'
GROUP BY
    l_returnflag;
"""


def midcard_frame(ns, table, key, engine=None):
    """The query bench.py's extra.midcard runs (7 aggregates over 6 accumulators) -- shared with the bench so that the port
    is pinned on exactly what it checks there."""
    p, q = ns.Col("l_extendedprice"), ns.Col("l_quantity")
    return ns.DataFrame(engine).table(str(table)).group_by(ns.Col(key)).agg(
        ns.F.count().alias("n"), ns.F.sum(q).alias("sum_q"), ns.F.sum(p).alias("sum_p"), ns.F.sum(p * q).alias("sum_pq"),
        ns.F.min(p).alias("min_p"), ns.F.max(p).alias("max_p"), ns.F.avg(p).alias("avg_p"))
