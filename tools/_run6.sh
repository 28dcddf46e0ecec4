python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/prof.sh r2h cfg5_chr1_10kb_band "estep_bulk_kernel|emit_kernel|quantise" 9 3
