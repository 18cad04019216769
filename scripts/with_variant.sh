#!/bin/bash
# usage: scripts/with_variant.sh build/variants/X.so <command...>: run a command with a prebuilt library variant in place
# of covid19_agentbasedsimulation_b200/libcovidabs.so (timing-only / diagnostic builds), then put the real one back
v="$1"; shift
L=covid19_agentbasedsimulation_b200/libcovidabs.so
cp "$L" /tmp/keep_lib.so
cp "$v" "$L"; touch "$L"
CABS_AB_OLD_LIB=1 "$@"; rc=$?
cp /tmp/keep_lib.so "$L"; touch "$L"
exit $rc
