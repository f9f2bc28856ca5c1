#!/bin/bash
echo "== pytest -m gpu (full suite, committed state)"
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 2>&1 | tail -16
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
echo done
