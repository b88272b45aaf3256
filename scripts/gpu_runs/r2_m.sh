# call M (last call of the round): smoke() with the sector engine in it, whole -m gpu suite on the final commit
mkdir -p gpurun_out
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r2m_smoke.txt; cat gpurun_out/r2m_smoke.txt
( timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2m_gputests.txt; cat gpurun_out/r2m_gputests.txt
