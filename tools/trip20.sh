#!/usr/bin/env bash
set -uo pipefail
echo "== full GPU suite"; timeout 1500 python -m pytest tests -m gpu -q -rs 2>&1 | tail -4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
