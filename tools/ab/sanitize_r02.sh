echo "compute-sanitizer on a B200 (gpurun), final round-2 libp25cu.so: cluster channelizer kernel (768 threads, st.async + mbarrier exchange,"
echo "pipelined batches), walker with the BF16 tensor-pipe sync prefilter and its idle loop, PDU receive, packed asynchronous poll."
echo "An earlier memcheck pass of this list FAILED test_packed_poll_matches_records_and_runs_async with 0 memory errors: under the tool's"
echo "timing the walker of chunk k+1 (ctx->stream) overtook the compaction of the poll started after chunk k (stream2) and the"
echo "events of chunk k were delivered twice.  Fixed in p25cu_decode (the walker now waits for the started poll's event)."
echo "== memcheck: pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -k 'channelizer or golden or prefilter or pdu or packed_poll or pinned'"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -k "channelizer or golden or prefilter or pdu or packed_poll or pinned" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|FAILED" | tail -6
echo "== racecheck: -k 'golden or prefilter or pdu or invariant_under_chunking'"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -k "golden or prefilter or pdu or invariant_under_chunking" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|FAILED" | tail -6
echo "== synccheck: -k 'golden or prefilter or invariant_under_chunking'"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -k "golden or prefilter or invariant_under_chunking" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|FAILED" | tail -5
echo "== memcheck: tests/ab/dbg_poll.py (records / packed synchronous / packed asynchronous against the oracle)"
compute-sanitizer --tool memcheck python tests/ab/dbg_poll.py 2>&1 | grep -E "differing|ERROR SUMMARY"
