set -x
timeout 230 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
