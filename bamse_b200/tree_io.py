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
    """one plaintext dot node per sample (reference tree_io.py:61-73)."""
    out = ''
    for n, name in enumerate(sample_names):
        shifted = [-1] + list((np.array(tree) + 1).astype(int))       # virtual germline root 0
        colors = tuple(["black"] + list(cluster_colors))
        vafs = [1.0] + list(vafs_tr[n])
        present = [0] + [x for x in range(1, len(vafs)) if (vafs[x] > 0.04)]
        for x in range(len(vafs)):
            if x not in present:
                shifted[x] = -5
        body = table_expression(height, width, border, 'v', shifted, 0, vafs, colors)[4:-5]
        out += name + ' [\n    shape=plaintext\n    xlabel ="' + name + '"\n    label=<\n' + body + '\n  >];'
    return out


def chart_plot(subclone_letter, color, percentages, sample_names, sample_colors, num_muts, bar_width=30, border=3):
    """bar chart node of one subclone (reference tree_io.py:96-115)."""
    out = ('Subclone' + subclone_letter + ' [\n      shape=plaintext\n      label=<<table FIXEDSIZE="TRUE" HEIGHT="170" WIDTH="'
           + str(bar_width * len(sample_names) + 60) + '" color="' + color + '" bgcolor="' + color + '" border="'
           + str(border) + '" cellborder="0" cellspacing="0">\n')
    out += '<tr><td rowspan ="3">' + subclone_letter + '-' + str(num_muts) + ' </td>\n'
    for n in range(len(sample_names)):
        pct = int(100 * percentages[n])
        if pct < 100:
            out += ('<td FIXEDSIZE="TRUE" HEIGHT="110" WIDTH="' + str(bar_width)
                    + '" cellpadding="0"><table cellspacing="0" cellpadding="0" cellborder="0" border="0"><tr><td '
                    + 'FIXEDSIZE="TRUE" HEIGHT="' + str(100 - pct) + '" WIDTH="' + str(bar_width - 2) + '" > </td></tr>')
        if pct > 2:
            out += (' <tr><td FIXEDSIZE="TRUE" HEIGHT="' + str(pct) + '" WIDTH="' + str(bar_width - 2) + '" bgcolor="'
                    + sample_colors[n] + '" > </td></tr>')
        out += '</table></td>'
    out += '</tr><tr>'
    for x in percentages:
        out += '<td>' + str(int(np.around(100 * x, 0))) + '</td>'
    out += '</tr><tr>'
    for x in sample_names:
        out += '<td>' + x + '</td>'
    return out + '</tr></table> \n>]'


def write_dot_files(dat, sample_colors, cluster_colors, treeoutfile, samplesoutfile):
    """graphviz files: per-sample tilings and the subclone tree (reference tree_io.py:77-93)."""
    body = put_sample_in_node(dat['tree'], dat['vaf'].transpose(), dat['sample_names'], sample_colors, cluster_colors,
                              height=150, width=150)
    with open(samplesoutfile, 'w') as f:
        f.write('digraph {\n' + body + '}\n')
    letters = [chr(65 + x) for x in range(len(dat['tree']))]
    body = ''
    for x in range(len(dat['tree'])):
        n_mut = len([y for y in dat['assign'] if y == x])
        body += chart_plot(letters[x], cluster_colors[x], dat['vaf'].transpose()[:, x], dat['sample_names'],
                           sample_colors, n_mut)
    body += '\n'
    for x in range(len(dat['tree'])):
        if not x == dat['root']:
            body += 'Subclone' + letters[int(dat['tree'][x])] + ' -> Subclone' + letters[x] + '\n'
    with open(treeoutfile, 'w') as f:
        f.write('digraph {\n' + body + '}\n')


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
    """text report of the ranked solutions (reference tree_io.py:126-140)."""
    with open(destfile, 'w') as f:
        for i, tree in enumerate(data):
            f.write('Solution Number ' + str(i + 1) + '\n')
            f.write('tree = ' + str(tree['tree']) + '\n')
            depth = get_tree_depth(tree['tree'])
            pad = max(depth) - np.array(depth) + 1
            labels = [chr(x + 65) + ' ' * 4 * pad[x] + str(np.around(tree['vaf'][x, :], 2))
                      for x in range(len(tree['tree']))]
            f.write(get_ascii_tree(tree['tree'], labels) + '\n')
            f.write('logscore = ' + str(tree['totalscore']) + '\n')
            f.write('ML Subclone Fractions = \n' + str(tree['clone_proportions']) + '\n')
            f.write('Mutations Subclone Membership = ' + str([chr(x + 65) for x in tree['assign']]) + '\n')
            f.write('----------------------------------------------------------------')
    return
