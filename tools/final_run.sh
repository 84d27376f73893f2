#!/bin/bash
# End-of-round evidence run on one B200 (through gpurun): GPU parity suite, smoke, the default bench line, the
# reference arm, the ncu launch list of the bench command and one `ncu --set full` capture of the dominant kernel on the
# speech-like input.  Everything lands in gpurun_out/<tag>_*; summaries are copied to profiles/ by hand afterwards.
#   gpurun --timeout 900 -- 'bash tools/final_run.sh r02f'
tag=${1:-final}
o=gpurun_out
mkdir -p $o
cd "$(dirname "$0")/.."
timeout 420 python -m pytest tests -m gpu -x -q > $o/${tag}_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 $o/${tag}_gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $o/${tag}_smoke.log
timeout 300 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_n1.json 2> $o/${tag}_bench_reference_n1.err; echo "reference arm exit $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-cross-check > $o/${tag}_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:logmel_tc_kernel -s 3 -c 1 -f -o $o/${tag}_tc_full_speech \
  python tools/time_kernel.py 256 128 3 bursty > $o/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
for k in noise bursty ragged; do timeout 60 python tools/time_kernel.py 256 128 20 $k 2>&1 | tail -1; done
head -c 1500 $o/${tag}_bench_n1.json
