#!/bin/bash
# Round-2 probe: is a real parasail reachable on the GPU pod? (VERDICT "Next round" item 1)
out=gpurun_out/parasail_probe_r02.log
{
echo "== date"; date -u
echo "== import parasail"; python -c "import parasail; print(parasail.__version__, parasail.__file__)" 2>&1 | tail -2
echo "== pip download"; timeout 30 python -m pip download parasail==1.3.4 --no-deps -d /tmp/pd 2>&1 | tail -3
echo "== pip wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "parasail|biopython|spoa|snapgene|pulp|pysam"; echo "(end wheelhouse grep)"
echo "== pip install from wheelhouse"; python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/pt parasail 2>&1 | tail -2
echo "== baseline/_ref"; ls -la baseline/_ref 2>&1 | head
echo "== find libparasail"; find / \( -iname 'libparasail*' -o -iname 'parasail*.so' -o -iname 'parasail-*' -o -iname 'parasail_*' -o -iname 'parasail.h' \) -not -path '/proc/*' 2>/dev/null | head
echo "== find dirs named parasail"; find / -type d -iname '*parasail*' -not -path '/proc/*' 2>/dev/null | head
echo "== network"; timeout 8 python - <<'PY'
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=5); print("pypi reachable")
except Exception as e: print("pypi unreachable:", e)
PY
echo "== nproc"; nproc; free -g | head -2
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
} > $out 2>&1
cat $out
