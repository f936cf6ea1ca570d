"""Drop-in command line of BAMSE on the B200 scorer.

Same interface as the reference's bamse.py (argparse block bamse.py:29-36 kept verbatim:
metavars, defaults, help strings, `inputfile` nargs='+' using the first) and the same output
files in the current directory (bamse.py:115-118): tree_number_<t>_subclones.dot,
tree_number_<t>_samples.dot, tree_number_<t>.txt, bamse_output.txt.  Input parsing, KMeans
candidate generation and the output writers stay on the host; the scoring loop
bamse.py:91-106 is one batched call into libbamse_cuda (analyse_clusterings).

`-s` is parsed and carried but, exactly as in the reference source (bamse.py:33,40,100),
does not influence any score.
"""
import argparse
import re
import sys

import numpy as np

cluster_colors = ["plum", "lightyellow", "palegreen", "pink", "paleturquoise", "tan", "moccasin", "lightcyan3",
                  "oldlace", "palevioletred", "olivedrab1", "lightgray", "steelblue1"]
sample_colors = ["red", "firebrick4", "brown4", "darkgoldenrod4", "deeppink1", "darkgreen", "cyan4", "darkorchid",
                 "blue", "dimgray"]


def build_parser():
    parser = argparse.ArgumentParser(description='Generates likely tumor evolution histories given somatic readcounts.')
    parser.add_argument('infile', metavar='inputfile', type=str, nargs='+', help='an integer for the accumulator')
    parser.add_argument('-e', metavar='error', default=0.001, type=float, help='Sequencing error probability (Default:0.001)')
    parser.add_argument('-nc', metavar='max_clusts', default=12, type=int, help='Maximum possible number of subclones')
    parser.add_argument('-s', metavar='sparsity', default=-1, type=float, help='this controls sparsity of solutions, between 0. and .1, considered 0 otherwise')
    parser.add_argument('-n', metavar='n_trees', default=3, type=int, help='Number of clusterings candidates analysed per number of clusters')
    parser.add_argument('-t', metavar='top_trees', default=1, type=int, help='Number of top solutions to return')
    # additions (defaults reproduce the reference behaviour)
    parser.add_argument('--device', default=0, type=int, help='CUDA device ordinal (B200)')
    parser.add_argument('--gpus', default=1, type=int, help='number of GPUs of this box to use (devices --device .. --device+N-1; one process drives them all)')
    parser.add_argument('--check', action='store_true', help='score everything in the fp64 check mode')
    parser.add_argument('--seed', default=None, type=int, help='seed numpy/sklearn RNG for reproducible KMeans restarts')
    parser.add_argument('--tree-budget', default=2e11, type=float,
                        help='refuse to start when the clusterings to analyse span more labeled rooted trees than '
                             'this (the search is exhaustive: K^(K-1) trees per K-cluster clustering; default 2e11)')
    return parser


def _fmt_duration(sec):
    if sec < 120:
        return "%.0f s" % sec
    if sec < 7200:
        return "%.0f min" % (sec / 60.0)
    return "%.1f h" % (sec / 3600.0)


def read_input(path):
    """bamse.py:53,64-72: tab-separated table with <sample>.ref / <sample>.var columns."""
    import pandas as pd
    table_data = pd.read_table(path)
    cols = table_data.columns.values
    ref_names = [m.string.replace('.ref', '') for m in filter(None, [re.match('(.*)?.ref', x) for x in cols])]
    var_names = [m.string.replace('.var', '') for m in filter(None, [re.match('(.*)?.var', x) for x in cols])]
    sample_names = sorted(set.intersection(set(ref_names), set(var_names)))
    As = table_data[[x + '.var' for x in sample_names]].values
    Bs = table_data[[x + '.ref' for x in sample_names]].values
    return As, Bs, sample_names


def main(argv=None):
    args = build_parser().parse_args(argv)
    from . import cuda_shim
    from .clustering import analyse_clusterings, get_context
    from .tree_io import write_dot_files, write_text_output
    from .tree_tools import remove_zeros
    from .workloads import candidate_clusterings

    inputfile = args.infile[0]
    err, sparsity, n_trees, max_clusts, top_trees = args.e, args.s, args.n, args.nc, args.t  # noqa: F841 (sparsity unused, as upstream)
    if args.seed is not None:
        np.random.seed(args.seed)
    print(max_clusts)
    X_As, X_Bs, sample_names = read_input(inputfile)
    print(sample_names)
    print(X_As.shape)
    if args.gpus > 1:
        from .multigpu import MultiContext
        ctx = MultiContext(range(args.device, args.device + args.gpus))
    else:
        ctx = get_context(args.device)
    clusts, pf = candidate_clusterings(X_As, X_Bs, max_clusts, n_trees, err, ctx=ctx, verbose=True)
    if any(int(c.As.shape[0]) > cuda_shim.KMAX for c in clusts):
        raise SystemExit("bamse: more than %d clusters is not supported by the CUDA scorer" % cuda_shim.KMAX)
    # The reference prunes with a heuristic branch-and-bound; this scorer searches all K^(K-1) labeled rooted
    # trees of every clustering exactly.  Say what that means before starting (K = 11: 2.6e10, K = 12: 7.4e11).
    n_trees_total = sum(int(c.As.shape[0]) ** (int(c.As.shape[0]) - 1) for c in clusts)
    # K <= 10 is table driven (one windowed dot product per tree and sample: 1e8 - 1e9 trees/s per B200, measured on
    # the BASELINE configurations); K = 11, 12 still cost about one convolution per tree and sample (~2e7 trees/s)
    n_small = sum(int(c.As.shape[0]) ** (int(c.As.shape[0]) - 1) for c in clusts if int(c.As.shape[0]) <= 10)
    eta = (n_small / 2.0e8 + (n_trees_total - n_small) / 2.0e7) / max(1, args.gpus)
    print('analysing %d clusterings on %d GPU(s): %.3g labeled rooted trees, roughly %s (exhaustive and exact; the '
          'reference prunes heuristically)' % (len(clusts), max(1, args.gpus), n_trees_total, _fmt_duration(eta)))
    if n_trees_total > args.tree_budget:
        raise SystemExit("bamse: %.3g trees exceed --tree-budget %.3g (clusterings with K >= 11 take hours per "
                         "clustering); lower -nc / -n or raise --tree-budget explicitly" % (n_trees_total, args.tree_budget))
    trees = analyse_clusterings(clusts, top_trees, ctx=ctx,
                                precision=cuda_shim.FP64 if args.check else cuda_shim.FP32)
    for x in trees:
        x['sample_names'] = list(sample_names)
    for t in range(min(top_trees, len(trees))):
        this_tree = remove_zeros(trees[t])
        this_tree['vaf'] = this_tree['Bmatrix'].dot(this_tree['clone_proportions'])
        write_dot_files(this_tree, sample_colors, cluster_colors, 'tree_number_' + str(t) + '_subclones.dot',
                        'tree_number_' + str(t) + '_samples.dot')
        write_text_output([this_tree], 'tree_number_' + str(t) + '.txt')
    write_text_output(trees[:min(top_trees, len(trees))], 'bamse_output.txt')
    return 0


if __name__ == "__main__":
    sys.exit(main())
