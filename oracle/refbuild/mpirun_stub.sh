#!/bin/bash
# mpirun stand-in for the MPI stub (oracle/refbuild/mpistub): starts P copies of a command as separate processes
# with MPISTUB_RANK / MPISTUB_SIZE / MPISTUB_DIR set and waits for all of them.  Test infrastructure only.
#   mpirun_stub.sh P command [args...]
set -u
P=$1; shift
DIR=$(mktemp -d /dev/shm/mpistub.XXXXXX 2>/dev/null || mktemp -d /tmp/mpistub.XXXXXX)
pids=()
for ((r = 0; r < P; ++r)); do
  if [ "$r" -eq 0 ]; then
    MPISTUB_RANK=$r MPISTUB_SIZE=$P MPISTUB_DIR=$DIR "$@" &
  else
    MPISTUB_RANK=$r MPISTUB_SIZE=$P MPISTUB_DIR=$DIR "$@" > "$DIR/rank$r.log" 2>&1 &
  fi
  pids+=($!)
done
rc=0
for ((r = 0; r < P; ++r)); do
  wait "${pids[$r]}" || { rc=$?; echo "mpirun_stub: rank $r exited with $rc" >&2; [ "$r" -gt 0 ] && tail -5 "$DIR/rank$r.log" >&2; }
done
rm -rf "$DIR"
exit $rc
