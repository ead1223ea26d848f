#!/bin/bash
# Same-box A/B of two builds of libmreserve_b200.so: tools/ab_lib.sh <other.so> <command...>
# runs the command with the in-tree library, then with <other.so> swapped in, then restores.
set -u
OTHER="$1"; shift
LIB=merlot_reserve_b200/lib/libmreserve_b200.so
cp "$LIB" /tmp/_ab_new.so
echo "== NEW"; "$@"
cp "$OTHER" "$LIB"
echo "== OTHER ($OTHER)"; "$@"
cp /tmp/_ab_new.so "$LIB"
echo "== NEW (again)"; "$@"
