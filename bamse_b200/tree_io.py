"""Output writers, byte-compatible with the reference's tree_io.py for the same tree dicts
(keys: tree, root, assign, vaf, clone_proportions, totalscore, sample_names).  Host-side
string formatting only.  The ASCII tree drawing restates asciitree.LeftAligned with its
default BoxStyle (the `asciitree` package is not installed in this image; parity of that one
piece is pinned by construction rules, not by running the package).
"""
import numpy as np

from .tree_tools import get_childs, get_tree_depth


# ------------------------------------------------------------------ sample tiling (dot/html)
def _cell(height, width, border, color, bgcolor):
    return ('<td FIXEDSIZE="TRUE" HEIGHT="%s" WIDTH="%s" border="%s" color="%s" bgcolor="%s"></td>\n'
            % (height, width, border, color, bgcolor))


def _blank(height, width, border):
    return '<td FIXEDSIZE="TRUE" HEIGHT="%s" WIDTH="%s" border="%s" color=\'white\'></td>' % (height, width, border)


def table_expression(remheight, remwidth, border, orient, tree, node, vafs, colors):
    """nested html table tiling one sample's subclones (reference tree_io.py:7-58)."""
    bg = ['white'] + list(colors[1:])
    kids = get_childs(tree, node)
    if not kids:
        return _cell(remheight, remwidth, border, colors[node], bg[node])
    out = ('<td><table  FIXEDSIZE="TRUE" HEIGHT="%s" WIDTH="%s" bgcolor="%s" border="%s" cellborder=\'0\' color="%s"'
           ' cellspacing=\'0\' cellpadding=\'0\'>\n' % (remheight, remwidth, bg[node], border, colors[node]))
    own = vafs[node] - sum([vafs[k] for k in kids])
    shares = np.array([0.0, own] + [vafs[k] for k in kids])
    shares = shares / sum(shares)
    cum = np.cumsum(shares)
    if orient == 'h':
        out += '<tr>'
        w_eff = remwidth - border * (len(kids) + 1)
        h_eff = remheight - 2 * border
        widths = np.diff(np.around(w_eff * cum, 0)).astype(int)
        if widths[0] > 2 * border:
            out += _blank(h_eff, widths[0], border) + '\n'
        else:
            widths[1] += border
        for n, k in enumerate(kids):
            out += table_expression(h_eff, widths[n + 1], border, 'v', tree, k, vafs, colors)
        out += '\n</tr>'
    if orient == 'v':
        h_eff = remheight - border * (len(kids) + 1)
        w_eff = remwidth - 2 * border
        heights = np.diff(np.around(h_eff * cum, 0)).astype(int)
        if heights[0] > 2 * border:
            out += '<tr>' + _blank(heights[0], w_eff, border) + '\n</tr>'
        else:
            heights[1] += border
        for n, k in enumerate(kids):
            out += '<tr>' + table_expression(heights[n + 1], w_eff, border, 'h', tree, k, vafs, colors) + '</tr>\n'
        out += '\n'
    return out + '</table></td>'


def put_sample_in_node(tree, vafs_tr, sample_names, sample_colors, cluster_colors, height=100, width=100, border=2,
                       orient='v'):
    """one plaintext dot node per sample (reference tree_io.py:61-73): the subclone tree is hung under
    a virtual germline node 0; subclones below 4 % prevalence in the sample are cut off."""
    palette = ("black",) + tuple(cluster_colors)
    nodes = []
    for name, row in zip(sample_names, vafs_tr):
        prevalence = [1.0] + list(row)
        parents = [-1] + [int(p) + 1 for p in np.array(tree)]
        for k in range(1, len(prevalence)):
            if not prevalence[k] > 0.04:
                parents[k] = -5
        tiling = table_expression(height, width, border, 'v', parents, 0, prevalence, palette)
        nodes.append('%s [\n    shape=plaintext\n    xlabel ="%s"\n    label=<\n%s\n  >];' % (name, name, tiling[4:-5]))
    return ''.join(nodes)


_BAR_EMPTY = ('<td FIXEDSIZE="TRUE" HEIGHT="110" WIDTH="%d" cellpadding="0"><table cellspacing="0" cellpadding="0" '
              'cellborder="0" border="0"><tr><td FIXEDSIZE="TRUE" HEIGHT="%d" WIDTH="%d" > </td></tr>')
_BAR_FILL = ' <tr><td FIXEDSIZE="TRUE" HEIGHT="%d" WIDTH="%d" bgcolor="%s" > </td></tr>'


def chart_plot(subclone_letter, color, percentages, sample_names, sample_colors, num_muts, bar_width=30, border=3):
    """bar chart node of one subclone: one bar per sample, the rounded percentages, the sample names
    (reference tree_io.py:96-115)."""
    parts = ['Subclone%s [\n      shape=plaintext\n      label=<<table FIXEDSIZE="TRUE" HEIGHT="170" WIDTH="%d" color="%s" '
             'bgcolor="%s" border="%d" cellborder="0" cellspacing="0">\n'
             % (subclone_letter, bar_width * len(sample_names) + 60, color, color, border),
             '<tr><td rowspan ="3">%s-%d </td>\n' % (subclone_letter, num_muts)]
    for n in range(len(sample_names)):
        filled = int(100 * percentages[n])
        if filled < 100:
            parts.append(_BAR_EMPTY % (bar_width, 100 - filled, bar_width - 2))
        if filled > 2:
            parts.append(_BAR_FILL % (filled, bar_width - 2, sample_colors[n]))
        parts.append('</table></td>')
    parts.append('</tr><tr>')
    parts.extend('<td>%d</td>' % int(np.around(100 * x, 0)) for x in percentages)
    parts.append('</tr><tr>')
    parts.extend('<td>%s</td>' % x for x in sample_names)
    parts.append('</tr></table> \n>]')
    return ''.join(parts)


def write_dot_files(dat, sample_colors, cluster_colors, treeoutfile, samplesoutfile):
    """graphviz files: per-sample tilings and the subclone tree (reference tree_io.py:77-93)."""
    K = len(dat['tree'])
    per_sample = dat['vaf'].transpose()
    with open(samplesoutfile, 'w') as f:
        f.write('digraph {\n%s}\n' % put_sample_in_node(dat['tree'], per_sample, dat['sample_names'], sample_colors,
                                                         cluster_colors, height=150, width=150))
    letters = [chr(65 + k) for k in range(K)]
    sizes = [sum(1 for y in dat['assign'] if y == k) for k in range(K)]
    charts = [chart_plot(letters[k], cluster_colors[k], per_sample[:, k], dat['sample_names'], sample_colors, sizes[k])
              for k in range(K)]
    edges = ['Subclone%s -> Subclone%s\n' % (letters[int(dat['tree'][k])], letters[k]) for k in range(K)
             if not k == dat['root']]
    with open(treeoutfile, 'w') as f:
        f.write('digraph {\n%s\n%s}\n' % (''.join(charts), ''.join(edges)))


# ------------------------------------------------------------------------------ text report
class LeftAligned(object):
    """asciitree.LeftAligned with the default BoxStyle (ASCII glyphs, horiz_len 2, indent 1,
    label_space 1) over a {label: {child_label: {...}}} dict."""

    def _render(self, label, children):
        lines = [label]
        items = list(children.items())
        for n, (lab, sub) in enumerate(items):
            block = self._render(lab, sub)
            if n == len(items) - 1:
                lines.append(' +-- ' + block[0])            # last_child_head
                lines.extend('    ' + l for l in block[1:])       # last_child_tail: indent + ' ' + 2 spaces
            else:
                lines.append(' +-- ' + block[0])            # child_head: indent, '+', 2 x '-', label_space
                lines.extend(' |  ' + l for l in block[1:])       # child_tail: indent + '|' + 2 spaces
        return lines

    def __call__(self, tree):
        (label, children), = list(tree.items())[:1]
        return '\n'.join(self._render(label, children))


def get_asciitree_dict(tree, labels, node):
    return {labels[x]: get_asciitree_dict(tree, labels, x) for x in get_childs(tree, node)}


def get_ascii_tree(tree, labels):
    """reference tree_io.py:120-124."""
    root = list(tree).index(-1)
    return LeftAligned()({labels[root]: get_asciitree_dict(tree, labels, root)})


def write_text_output(data, destfile):
    """text report of the ranked solutions (reference tree_io.py:126-140): parent vector, ASCII tree
    annotated with the per-sample prevalences, total log score, ML fractions, memberships."""
    rule = '-' * 64
    with open(destfile, 'w') as f:
        for rank, sol in enumerate(data, start=1):
            parents = sol['tree']
            depth = np.array(get_tree_depth(parents))
            indent = depth.max() - depth + 1
            labels = ['%s%s%s' % (chr(65 + k), ' ' * (4 * int(indent[k])), np.around(sol['vaf'][k, :], 2))
                      for k in range(len(parents))]
            f.write('Solution Number %d\n' % rank)
            f.write('tree = %s\n' % (parents,))
            f.write('%s\n' % get_ascii_tree(parents, labels))
            f.write('logscore = %s\n' % (sol['totalscore'],))
            f.write('ML Subclone Fractions = \n%s\n' % (sol['clone_proportions'],))
            f.write('Mutations Subclone Membership = %s\n' % ([chr(65 + int(x)) for x in sol['assign']],))
            f.write(rule)
    return
