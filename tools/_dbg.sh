echo "=== sanitizer c1"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/exp_traverse_noprof.py c1 1 2>&1 | grep -v "^$" | head -50
echo "=== debug c2"; timeout 600 python tools/debug_pixels.py c2 4 2>&1 | tail -40
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_baseline_sizes_gpu.py 2>&1 | tail -12
