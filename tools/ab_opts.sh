# A/B of context options on the headline frame (and slices): bash tools/ab_opts.sh "opt=val,opt=val" "..." ; PROBE_CASES from the environment
for o in "$@"; do echo "== $o"; PROBE_OPTS=$o python tools/probe_slice.py; done
