export PROBE_CASES="1,0;4,3;8,3;8,0"
for o in "help_window=0" "help_window=8" "help_window=32" "help_window=64" "help_window=256" "help_window=64,spp_chunks=2" "help_window=0,spp_chunks=2" "help_window=0"; do echo "== $o"; PROBE_OPTS=$o python tools/probe_slice.py; done
