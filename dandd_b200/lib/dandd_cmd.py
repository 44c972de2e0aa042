"""Command-line front end: `dandd tree | progressive | kij` (reference lib/dandd_cmd.py).

Same sub-commands, flags, defaults and output file names as the reference, so existing scripts
and pickles keep working.  The reference registers an `info` sub-parser whose handler is commented
out (reference :234-246), so `dandd info` dies there with an AttributeError; here the same flags are
accepted and the command does what its help text says: it loads a tree pickle, adds the sketches a
--ksweep asks for, prints the tree's delta table and (with --ksweep) writes the per-k table of every
node.  Additions: `--device N` selects the GPU (default: LOCAL_RANK or 0), `--gpus N`.  Flag tables below
cite the reference lines they mirror.
"""
import argparse
import os
import pickle
import sys

import huffman_dandd
from huffman_dandd import write_listdict_to_csv

VERSION = "%(prog)s 1.0.0 (dandd_b200)"
COMMAND_LINE = False   # set by the `dandd` launcher: process-wide environment tweaks are only made for a real command-line run


def insert_pre_ext(filename, string):
    stem, dot, ext = filename.rpartition(".")
    return f"{stem}.{string}.{ext}" if dot else f"{string}.{filename}"


# (flags, keyword arguments) tables -------------------------------------------------------------------
UNIVERSAL = [  # reference :23-30
    (("--version",), dict(action="version", version=VERSION)),
    (("--verbose", "-v"), dict(action="store_true", default=False, help="Print some trees and report steps of actions.")),
    (("--debug",), dict(action="store_true", default=False, dest="debug", help="Show the commands this run stands in for.")),
    (("--lowmem",), dict(action="store_true", default=False, dest="lowmem",
                         help="Trust stored cardinalities of multi-fasta sketches whose files are gone. Not with --safe.")),
    (("--safe",), dict(action="store_true", default=False, dest="safety",
                       help="Re-verify every sketch name hash against the sum of its component fasta hashes.")),
    (("--fast",), dict(action="store_true", default=False, dest="fast", help="Don't save so much stuff for second usage.")),
    (("--device",), dict(type=int, default=None, dest="device", metavar="GPU", help="CUDA device to use (default LOCAL_RANK or 0).")),
    (("--gpus",), dict(type=int, default=None, dest="gpus", metavar="N",
                       help="tree: use N GPUs of this node, one process each (what `torchrun --nproc-per-node N` sets up, "
                            "without the launcher's start-up cost). N is an upper bound: one process per "
                            "DANDD_B200_BYTES_PER_GPU (default 12 GiB; 0 = all N if any) of FASTA the sketch database has never "
                            "seen, all N with --exact.")),
]
KSWEEP = [  # reference :145-150
    (("--ksweep",), dict(dest="ksweep", default=None, action="store_true",
                         help="sweep k for every combination; without --mink/--maxk the range is 2..32")),
    (("--mink",), dict(dest="mink", metavar="MINIMUM-K", required=False, default=2, type=int, help="smallest k of the sweep")),
    (("--maxk",), dict(dest="maxk", metavar="MAXIMUM-K", required=False, default=32, type=int, help="largest k of the sweep")),
]
TREE = [  # reference :166-200
    (("-s", "--tag"), dict(dest="tag", metavar="PREFIX TAG", type=str, required=False, default="dandd",
                           help="tag used to label output files")),
    (("-x", "--exact"), dict(dest="exact", default=False, action="store_true", required=False,
                             help="count k-mers exactly (KMC semantics) instead of estimating")),
    (("-d", "--datadir"), dict(dest="genomedir", default=None, type=str, metavar="FASTADIR",
                               help="directory of fasta files; all are used unless --fastas is given")),
    (("-o", "--out"), dict(dest="outdir", default=os.getcwd(), type=str, metavar="OUTPUT DIR", help="output directory")),
    (("-c", "--sketchdir"), dict(dest="sketchdir", default=None, type=str, metavar="SKETCHDIR",
                                 help="sketch database directory (default <out>/sketchdb)")),
    (("-k", "--kstart"), dict(dest="kstart", default=12, type=int, metavar="KSTART", help="k at which the search for delta starts")),
    (("-f", "--fastas"), dict(dest="flist_loc", metavar="FILEPATH", type=str, default=None,
                              help="file listing the fasta paths to use, one per line")),
    (("-l", "--label"), dict(dest="label", metavar="SUFFIX TAG", default="", required=False, help="extra label in output names")),
    (("-n", "--nchildren"), dict(dest="nchildren", metavar="INTEGER", type=int, default=None,
                                 help="children per tree node (default: all leaves under one root)")),
    (("-r", "--registers"), dict(dest="registers", metavar="INTEGER", default=20, help="log2 of the number of HLL registers")),
    (("-e", "--nthreads"), dict(dest="nthreads", metavar="INTEGER", type=int, default=0,
                                help="kept for compatibility (KMC thread count in the reference); unused on the GPU")),
    (("-C", "--no-canon"), dict(action="store_false", default=True, dest="canonicalize", help="use non-canonical k-mers")),
]
PROGRESSIVE = [  # reference :211-228
    (("-d", "--dtree"), dict(dest="delta_tree", metavar="DELTA TREE", required=True, help="pickle produced by the tree command")),
    (("-s", "--tag"), dict(dest="tag", metavar="species/experiment-tag-string", type=str, required=False, help="output tag")),
    (("-r", "--orderings"), dict(dest="ordering_file", metavar="ORDERING PICKLE", type=str, default=None,
                                 help="pickle of orderings, if not the default one named after the tag")),
    (("-f", "--fastas"), dict(dest="flist_loc", default=None, type=str, metavar="FILEPATH", help="subset (and order) of fastas")),
    (("-n", "--norderings"), dict(dest="norderings", default=0, type=int, metavar="NUM", help="number of random orderings")),
    (("-o", "--outdir"), dict(dest="outdir", default=os.getcwd(), type=str, metavar="OUTPUT DIR", help="output directory")),
    (("-l", "--label"), dict(dest="label", metavar="SUFFIX TAG", default="", required=False, help="extra label in output names")),
    (("--step",), dict(dest="step", default=1, type=int, metavar="INTEGER", help="fastas added per progression step")),
]
KIJ = [  # reference :260-272
    (("-d", "--dtree"), dict(dest="delta_tree", metavar="DELTA TREE", required=True, help="pickle produced by the tree command")),
    (("-s", "--tag"), dict(dest="tag", metavar="PREFIX TAG", type=str, required=False, help="output tag")),
    (("-f", "--fastas"), dict(dest="flist_loc", default=None, type=str, metavar="FILEPATH", help="subset of fastas to compare")),
    (("-o", "--outdir"), dict(dest="outdir", default=os.getcwd(), type=str, metavar="OUTPUT DIR", help="output directory")),
    (("-l", "--label"), dict(dest="label", metavar="SUFFIX TAG", default="", required=False, help="extra label in output names")),
    (("--afproject",), dict(dest="afproject", default=False, action="store_true", help="also write the AFproject tuple pickle")),
    (("--jaccard",), dict(dest="jaccard", default=False, action="store_true", help="also report per-k Jaccard")),
]


INFO = [  # reference :234-244
    (("-d", "--dtree"), dict(dest="delta_tree", metavar="DELTA TREE", required=True,
                             help="pickle produced by the tree command; nodes are updated to hold the sketches a --ksweep needs")),
    (("-s", "--tag"), dict(dest="tag", metavar="PREFIX TAG", type=str, required=False, help="output tag")),
    (("-o", "--outdir"), dict(dest="outdir", default=os.getcwd(), type=str, metavar="OUTPUT DIR", help="output directory")),
    (("-l", "--label"), dict(dest="label", metavar="SUFFIX TAG", default="", required=False, help="extra label in output names")),
]


def _add(parser, table):
    for flags, kwargs in table:
        parser.add_argument(*flags, **kwargs)
    return parser


def add_universal_cmds(subparser: argparse.ArgumentParser):
    return _add(subparser, UNIVERSAL)


def _select_device(args):
    if getattr(args, "device", None) is not None:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise SystemExit("--device selects the GPU of a single-process run; under torchrun / --gpus each rank uses LOCAL_RANK")
        os.environ["LOCAL_RANK"] = str(args.device)
    _narrow_visible_devices()


def _narrow_visible_devices() -> None:
    """A single-process run touches one GPU; on an 8-GPU node CUDA start-up is several seconds faster
    when only that one is visible (the driver otherwise sets up all eight).  Must run before anything
    initialises CUDA; does nothing under torchrun or when the user has set CUDA_VISIBLE_DEVICES."""
    if not COMMAND_LINE or int(os.environ.get("WORLD_SIZE", "1")) > 1 or "CUDA_VISIBLE_DEVICES" in os.environ:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("LOCAL_RANK", "0")
    os.environ["LOCAL_RANK"] = "0"


def _self_launch(args) -> list:
    """--gpus N outside torchrun: this process becomes rank 0 and starts ranks 1..N-1 as copies of its
    own command line, with the environment torchrun would have given them (RANK, LOCAL_RANK,
    WORLD_SIZE, MASTER_ADDR, MASTER_PORT on the loopback interface).  Returns the child processes."""
    import socket
    import subprocess
    n = int(args.gpus or 1)
    if n < 2 or "WORLD_SIZE" in os.environ:
        return []
    if not args.exact:
        # N is an upper bound.  Every further process costs about a second of (serialised) CUDA context
        # creation on top of its own interpreter start, which is what one GPU needs to pack and sketch
        # ~13 GB of FASTA for all k; and the part of a fresh run that does not shrink with N -- naming the
        # files, blake2b on the host's cores -- is the same however many processes share it.  So a worker
        # is started per BYTES_PER_GPU of FASTA the sketch database has never seen (measured: 8 x 3.1 GB
        # takes 8.8 s in one process on one GPU, 16-17 s in eight on eight), none for a re-run over named
        # files (what is left then is unions and the odd k outside the stored range).
        # DANDD_B200_BYTES_PER_GPU=0 starts all N whenever anything is fresh.
        per_gpu = int(float(os.environ.get("DANDD_B200_BYTES_PER_GPU", str(12 << 30))))
        fresh_bytes = sum(os.path.getsize(f) for f in _fresh_fastas(args))
        if fresh_bytes == 0:
            return []
        if per_gpu > 0:
            n = min(n, max(1, fresh_bytes // per_gpu))
            if n < 2:
                return []
    with socket.socket() as sock:       # a free port for the rendezvous store
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    base = dict(os.environ, WORLD_SIZE=str(n), LOCAL_WORLD_SIZE=str(n), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    base.setdefault("OMP_NUM_THREADS", "1")     # N processes share the host's cores (torchrun does the same)
    # (DANDD_B200_NARROW_DEVICES=1: each process sees only its own GPU, as device 0)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    ids = visible.split(",") if visible else [str(i) for i in range(n)]
    if len(ids) < n:
        raise SystemExit(f"--gpus {n}: only {len(ids)} device(s) in CUDA_VISIBLE_DEVICES")
    narrow = os.environ.get("DANDD_B200_NARROW_DEVICES", "0") != "0"   # measured: concurrent CUDA start-up is SLOWER with one visible device per process (4 ranks: 6.0 vs 3.7 s)

    def rank_env(r):
        return dict(RANK=str(r), LOCAL_RANK="0", CUDA_VISIBLE_DEVICES=ids[r]) if narrow else dict(RANK=str(r), LOCAL_RANK=str(r))
    os.environ.update(base, **rank_env(0))
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dandd")
    children = []
    for r in range(1, n):
        children.append(subprocess.Popen([sys.executable, script] + sys.argv[1:], env=dict(base, **rank_env(r))))
    return children


def _reap(children) -> None:
    bad = [c.args for c in children if c.wait() != 0]
    if bad:
        raise SystemExit(f"dandd: {len(bad)} worker process(es) failed")


def _single_rank_command(name) -> bool:
    """`progressive` and `kij` run on one GPU (their device work is milliseconds once the leaves are
    in HBM).  Started under torchrun, every rank but the first steps aside instead of repeating
    the job and racing on the output files.  Returns True if this process should return."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("RANK", "0")) > 0:
        print(f"dandd {name}: rank {os.environ.get('RANK')} idle (single-GPU command)")
        return True
    return False


def _load_tree(path):
    with open(path, "rb") as fh:
        return pickle.load(fh)


def _read_list(path):
    with open(path) as fh:
        return [line.strip() for line in fh]


def tree_command(args):
    """reference :43-62"""
    if not (args.genomedir or args.flist_loc):
        print("ERROR: You must provide either a datadirectory or a fasta file list!")
        sys.exit(1)
    children = _self_launch(args)       # --gpus N: this process is rank 0, the others start now
    _select_device(args)
    if not args.sketchdir:
        args.sketchdir = os.path.join(args.outdir, "sketchdb")
    from dandd_b200 import timing
    fresh = _early_prefetch(args)   # file reads + blake2b start now, under the ~4 s of torch import / CUDA start-up
    timing.mark("prefetch_started")
    if COMMAND_LINE and fresh and not args.exact:
        _start_engine_in_background()      # there is something to sketch: CUDA start-up (seconds) runs beside the
                                           # process-group rendezvous (a fully cached run never imports torch)
    with timing.span("init_ranks"):
        rank, world = _init_ranks()
    os.makedirs(args.sketchdir, exist_ok=True)
    os.makedirs(args.outdir, exist_ok=True)
    tool = "dashing"
    if args.exact:
        tool, args.registers = "kmc", 20
    if args.ksweep:
        args.ksweep = (int(args.mink), int(args.maxk))
    try:
        timing.mark("tree_start")
        dtree = huffman_dandd.create_delta_tree(
            tag=args.tag, genomedir=args.genomedir, sketchdir=args.sketchdir, kstart=args.kstart, nchildren=args.nchildren,
            registers=args.registers, flist_loc=args.flist_loc, canonicalize=args.canonicalize, tool=tool, debug=args.debug,
            nthreads=int(args.nthreads), safety=args.safety, fast=args.fast, verbose=args.verbose, ksweep=args.ksweep,
            lowmem=args.lowmem)
        timing.mark("tree_built")
        if dtree is not None:   # rank 0 (or the only process)
            with timing.span("save_outputs"):
                dtree.save(fileprefix=dtree.make_prefix(outdir=args.outdir, tag=args.tag, label=args.label), fast=args.fast)
    finally:
        if world > 1:           # release the ranks that serve exact-count requests, also when rank 0 fails
            store_mod = sys.modules.get("dandd_b200.store")          # only an ALREADY created store: creating one
            store = getattr(store_mod, "_store", None)               # here could mask the error being propagated
            workers = getattr(store, "exact_workers", None)
            if workers is not None:
                try:
                    workers.stop()
                except Exception as err:  # noqa: BLE001
                    print(f"dandd: could not release the exact-count workers: {err}", file=sys.stderr)
    with timing.span("finish_ranks"):
        _finish_ranks(world)
        _reap(children)


def _start_engine_in_background() -> None:
    import threading

    def start():
        try:
            huffman_dandd.get_store()
        except Exception:  # noqa: BLE001 -- the main thread will hit the same error and report it
            pass
    threading.Thread(target=start, name="dd-engine-start", daemon=True).start()


def _fresh_fastas(args) -> list:
    """The FASTAs of this run that the sketch database has never named (the whole job's, not one rank's).
    A fully cached re-run has none and reads nothing, like the reference (its fastahex pickle
    short-circuits the hash)."""
    try:
        fastas = huffman_dandd.list_fastas(args.genomedir, args.flist_loc)
    except (OSError, ValueError):
        return []               # create_delta_tree reports the problem
    sketchdir = args.sketchdir or os.path.join(args.outdir, "sketchdb")
    known = set()
    try:
        with open(os.path.join(sketchdir, "dandd_fastahex.pickle"), "rb") as fh:
            known = set(pickle.load(fh))
    except Exception:  # noqa: BLE001 -- no database yet (or unreadable): everything is new
        pass
    return [f for f in fastas if os.path.isfile(f) and os.path.basename(f) not in known]


def _early_prefetch(args) -> list:
    """Start reading (and hashing) this process's share of the FASTAs in background threads before
    anything imports torch.  Only files the sketch database has never seen.  Returns the files it
    started on."""
    from dandd_b200 import ingest
    from dandd_b200.shard import shard_by_size
    try:
        fastas = [f for f in huffman_dandd.list_fastas(args.genomedir, args.flist_loc) if os.path.isfile(f)]
    except (OSError, ValueError):
        return []
    fresh_all = set(_fresh_fastas(args))
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1 and len(fastas) >= world and not args.exact:
        owners = shard_by_size([os.path.getsize(f) for f in fastas], world)
        fastas = [fastas[i] for i in owners[rank]]
    elif world > 1 and rank > 0:
        return []               # split-genome and exact modes: rank 0 hashes, every rank reads on demand
    fresh = [f for f in fastas if f in fresh_all]
    ingest.prefetch(fresh)
    return fresh


def _init_ranks():
    """Under torchrun (WORLD_SIZE > 1) join the process group: nccl on GPUs, gloo otherwise."""
    if int(os.environ.get("WORLD_SIZE", "1")) < 2:
        return 0, 1
    from dandd_b200 import dist as dd_dist
    return dd_dist.init()


def _finish_ranks(world):
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()


def progressive_command(args):
    """reference :65-87"""
    if _single_rank_command("progressive"):
        return
    _select_device(args)
    dtree = _load_tree(args.delta_tree)
    args.tag = args.tag or dtree.speciesinfo.tag
    args.outfile = dtree.make_prefix(tag=args.tag, label=f"progu{args.norderings}", outdir=args.outdir)
    dtree.experiment.update(debug=args.debug, safety=args.safety, fast=args.fast, verbose=args.verbose, lowmem=args.lowmem,
                            baseset=set(), ksweep=(int(args.mink), int(args.maxk)) if args.ksweep else None)
    dtree.speciesinfo.update(tool=dtree.experiment["tool"])
    results, summary = dtree.progressive_wrapper(flist_loc=args.flist_loc, count=args.norderings,
                                                 ordering_file=args.ordering_file, step=args.step)
    write_listdict_to_csv(outfile=args.outfile + ".csv", listdict=results)
    write_listdict_to_csv(outfile=args.outfile + "summary.csv", listdict=summary)
    dtree.save(fileprefix=args.outfile)


def kij_command(args):
    """reference :107-132"""
    if _single_rank_command("kij"):
        return
    _select_device(args)
    dtree = _load_tree(args.delta_tree)
    dtree.speciesinfo.update(tool=dtree.experiment["tool"])
    args.tag = args.tag or dtree.speciesinfo.tag
    args.outfile = dtree.make_prefix(tag=args.tag, label=args.label, outdir=args.outdir)
    fastas = _read_list(args.flist_loc) if args.flist_loc else []
    if args.ksweep:
        dtree.experiment["ksweep"] = (int(args.mink), int(args.maxk))
    dtree.ksweep(mink=int(args.mink), maxk=int(args.maxk))
    kij_results, j_results = dtree.pairwise_spiders(sublist=fastas, mink=args.mink, maxk=args.maxk, jaccard=args.jaccard)
    write_listdict_to_csv(outfile=args.outfile + ".kij.csv", listdict=kij_results)
    if args.jaccard:
        write_listdict_to_csv(outfile=args.outfile + ".j.csv", listdict=j_results)
    dtree.speciesinfo.save_cardkey(dtree.experiment["tool"])
    dtree.speciesinfo.save_references(fast=False)
    if args.afproject:
        with open(args.outfile + "_AFtuples.pickle", "wb") as fh:
            pickle.dump(obj=dtree.prepare_AFproject(kij_results, j_results), file=fh)


def info_command(args):
    """`dandd info` (the reference's parser :234-246 has no handler): the delta table of a saved tree on
    stdout; with --ksweep every node is brought up to the k range first (sketching what is missing) and
    <prefix>_info.csv holds one row per (node, k) -- ngen, kval, card, delta_pos, title -- like the
    summary table of `progressive`."""
    if _single_rank_command("info"):
        return
    _select_device(args)
    dtree = _load_tree(args.delta_tree)
    dtree.speciesinfo.update(tool=dtree.experiment["tool"])
    args.tag = args.tag or dtree.speciesinfo.tag
    write_listdict_to_csv(outfile=None, listdict=dtree.report_deltas())
    if args.ksweep:
        lo, hi = int(args.mink), int(args.maxk)
        if dtree.experiment["tool"] == "dashing":
            hi = min(hi, huffman_dandd.HLL_MAX_K)
        dtree.experiment["ksweep"] = (lo, hi)
        dtree.ksweep(mink=lo, maxk=hi)
        rows = []
        for node in dtree._dt:
            rows.extend(node.summarize(mink=max(1, lo), maxk=hi))
        prefix = dtree.make_prefix(tag=args.tag, label=args.label, outdir=args.outdir)
        write_listdict_to_csv(outfile=prefix + "_info.csv", listdict=rows)
        dtree.speciesinfo.save_cardkey(dtree.experiment["tool"])
        dtree.speciesinfo.save_references(fast=False)


def parse_arguments():
    """Top-level parser + the list of sub-command names (reference :138-288)."""
    universal = add_universal_cmds(argparse.ArgumentParser(add_help=False))
    ksweep = _add(argparse.ArgumentParser(add_help=False), KSWEEP)
    parser = argparse.ArgumentParser(prog="DandD", description="program to explore delta values for a set of fasta files",
                                     parents=[universal])
    subparsers = parser.add_subparsers(title="subcommands", description="valid subcommands", help="additional help")
    subparsers.required = True
    specs = [
        ("tree", TREE, tree_command,
         "Calculate deltas for input fastas and full union. Create DandD tree object for further downstream analysis."),
        ("progressive", PROGRESSIVE, progressive_command,
         "Measure delta as each fasta is added to the set, over one given or several random orderings."),
        ("kij", KIJ, kij_command, "K Independent Jaccard (and optionally per-k Jaccard) for every pair of inputs."),
        ("info", INFO, info_command, "Print the delta table of a saved tree; with --ksweep also write every node's per-k table."),
    ]
    commands = []
    for name, table, handler, text in specs:
        sub = _add(subparsers.add_parser(name, help=text, parents=[universal, ksweep]), table)
        sub.set_defaults(func=handler)
        if name != "info":         # (the reference's list of command names leaves it out as well, :234)
            commands.append(name)
    return parser, commands
