#!/bin/bash
# yask.sh for the B200 engine: the launcher a user of the reference's src/kernel/yask.sh expects -- same option names for
# what still has a meaning on a GPU box, the same log-file convention and the same closing checks -- around
# yask_b200/bin/yask_kernel.<stencil>.b200.exe.  (The reference's script: /root/reference/src/kernel/yask.sh:240-330 options,
# :426-435 log name, :590-640 result checks.)  Written for this engine; nothing else of the reference's tooling is needed.
#
#   yask.sh -stencil <name> [-ranks <N>] [-log <file>] [-log_dir <dir>] [-exe <path>] [-exe_prefix <cmd>] [-pre_cmd <cmd>]
#           [-post_cmd <cmd>] [-v] [-dry_run] [-show_arch] [VAR=value ...] [--] [harness options, e.g. -g 1024 -trial_steps 50]
#
# -ranks N starts N processes (one per GPU of this node: RANK/WORLD_SIZE/LOCAL_RANK exported, no mpirun needed -- the ranks
# meet in the library's shared-memory mailbox); -v = a short validation run (-validate -trial_steps 4).
here=$(cd "$(dirname "$0")" && pwd)
bindir=$here/../bin
stencil="" arch=b200 nranks=1 logfile="" logdir=./logs exe="" exe_prefix="" pre_cmd=":" post_cmd=":" doval=0 dodry=0
envs=()
opts=()
invo="Script invocation: $0 $*"
while [[ $# -gt 0 ]]; do
    case "$1" in
        -h|-help) sed -n 2,14p "$0"; exit 0 ;;
        -show_arch) echo b200; exit 0 ;;
        -stencil) stencil=$2; shift 2 ;;
        -arch) arch=$2; shift 2 ;;
        -ranks) nranks=$2; shift 2 ;;
        -nodes|-host|-sh_prefix|-mpi_cmd) echo "note: option $1 has no effect (one node, no MPI launcher)"; shift 2 ;;
        -force_mpi|-offload) shift ;;
        -log) logfile=$2; shift 2 ;;
        -log_dir) logdir=$2; shift 2 ;;
        -exe) exe=$2; shift 2 ;;
        -exe_prefix) exe_prefix=$2; shift 2 ;;
        -pre_cmd) pre_cmd=$2; shift 2 ;;
        -post_cmd) post_cmd=$2; shift 2 ;;
        -v) doval=1; shift ;;
        -dry_run) dodry=1; shift ;;
        --) shift; opts+=("$@"); break ;;
        [A-Za-z_]*=*) envs+=("$1"); shift ;;
        *) opts+=("$1"); shift ;;
    esac
done
if [[ -z "$stencil" && -z "$exe" ]]; then echo "error: missing -stencil <name>" >&2; exit 1; fi
if [[ "$arch" != b200 ]]; then echo "error: this engine has one target, 'b200' (got -arch $arch)" >&2; exit 1; fi
: "${exe:=$bindir/yask_kernel.$stencil.$arch.exe}"
: "${logfile:=yask.$stencil.$arch.$(hostname).n1.r$nranks.$(date +%Y-%m-%d_%H-%M-%S)_p$$.log}"
[[ "$logfile" == */* ]] || logfile="$logdir/$logfile"
mkdir -p "$(dirname "$logfile")"
echo "Writing log to '$logfile'."
echo "$invo" > "$logfile"
if [[ ! -x "$exe" ]]; then echo "error: '$exe' not found or not executable." | tee -a "$logfile"; exit 1; fi
[[ $doval == 1 ]] && opts=(-validate -trial_steps 4 "${opts[@]}")
{
    echo "Num nodes: 1"; echo "Num ranks: $nranks"; echo "exe_prefix='$exe_prefix'"; echo "exe='$exe'"
    echo "pre_cmd='$pre_cmd'"; echo "post_cmd='$post_cmd'"
    command -v nvidia-smi > /dev/null && nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
} | tee -a "$logfile"
exe_str="$exe_prefix $exe ${opts[*]}"
echo "Binary invocation: ${envs[*]} $exe_str" | tee -a "$logfile"
if [[ $dodry == 1 ]]; then echo "YASK not started due to -dry_run option." | tee -a "$logfile"; echo "Log saved in '$logfile'."; exit 0; fi
echo "===================" | tee -a "$logfile"
(
    for e in "${envs[@]}"; do export "$e"; done
    sh -c "$pre_cmd"
    job="yask_sh_$$"
    pids=()
    for ((r = 1; r < nranks; r++)); do
        RANK=$r WORLD_SIZE=$nranks LOCAL_RANK=$r YASK_JOB_ID=$job $exe_prefix "$exe" "${opts[@]}" > "$logfile.rank$r" 2>&1 &
        pids+=($!)
    done
    RANK=0 WORLD_SIZE=$nranks LOCAL_RANK=0 YASK_JOB_ID=$job $exe_prefix "$exe" "${opts[@]}" 2>&1
    for p in "${pids[@]}"; do wait "$p" || echo "rank process $p failed"; done
    sh -c "$post_cmd"
) 2>&1 | tee -a "$logfile"
echo "===================" | tee -a "$logfile"
finish() { echo "Log saved in '$logfile'."; exit "$1"; }
if grep -q 'TEST FAILED' "$logfile" "$logfile".rank* 2> /dev/null; then echo "YASK did not pass internal validation test." | tee -a "$logfile"; finish 1; fi
if ! grep -q 'YASK DONE\|TEST PASSED' "$logfile"; then echo "YASK did not exit cleanly." | tee -a "$logfile"; finish 1; fi
grep -q 'TEST PASSED' "$logfile" && echo "YASK passed internal validation test." | tee -a "$logfile"
echo "YASK ran successfully." | tee -a "$logfile"
finish 0
