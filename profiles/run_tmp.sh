for i in 1 2 3 4 5 6; do timeout 600 python bench.py --no-cpu-baseline --no-breakdown --no-variants 2>/dev/null | python profiles/show_bench.py; done
