cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_r02.txt
: > $OUT
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool python tools/sanitize.py" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -v "^$" | tail -14 >> $OUT
done
cat $OUT
