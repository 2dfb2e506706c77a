"""Host-side mirror of the reference's ``OrgAlign`` result types.

The five-state DP itself (cost matrix, fill, traceback, landscape) runs in ``g2g_cost_dp``
(``csrc/g2g_dp.cu``) for all genes at once.  ``DP5`` and ``AlignmentLandscape`` here are per-gene
views with the reference's attribute names; the dense matrices are produced on first access by
re-running the kernel for that one gene in dump mode.
"""
import re

import numpy as np

from . import engine as _engine

STATES = "MWVDI"


class FiveStateMachine:
    """Transition probabilities / costs of the symmetric five-state machine
    (reference: OrgAlign.py:12-97).  Attribute names follow the reference (``P_xy`` = probability
    of moving to state x from state y, ``I_xy = -ln P_xy``)."""

    def __init__(self, P_mm, P_ii, P_mi, PROHIBIT_CASE=False):
        table = _engine.transition_costs((P_mm, P_ii, P_mi), PROHIBIT_CASE)
        self.table = table
        with np.errstate(over="ignore"):
            prob = np.exp(-table)
        for a, to in enumerate("mwvdi"):
            for b, frm in enumerate("mwvdi"):
                setattr(self, "I_%s%s" % (to, frm), table[a, b])
                setattr(self, "P_%s%s" % (to, frm), prob[a, b])
        # exact probabilities where exp(-(-ln p)) would round
        self.P_mm = self.P_ww = self.P_vv = P_mm
        self.P_ii = self.P_dd = P_ii
        self.P_mi = self.P_md = P_mi


class DP5:
    """Per-gene DP result view (reference: OrgAlign.py:206-646)."""

    def __init__(self, store, gene_idx, S, T, alignment_str, opt_cost, free_params):
        self._store = store
        self._gene_idx = gene_idx
        self.S = S
        self.T = T
        self.S_len = len(S.mean_trend)
        self.T_len = len(T.mean_trend)
        self.alignment_str = alignment_str
        self.opt_cost = opt_cost
        self.backward_run = False
        self._free_params = tuple(free_params)
        self._path = None
        self._dense = None
        self._fsa = None

    @property
    def FSA(self):
        if self._fsa is None:
            self._fsa = FiveStateMachine(*self._free_params, PROHIBIT_CASE=False)
        return self._fsa

    @property
    def alignment_path(self):
        """Cells visited by the traceback, end cell first, (0,0) excluded (OrgAlign.py:583-607);
        a mutable list, as downstream code appends to it."""
        if self._path is None:
            self._path = path_from_string(self.alignment_str)
        return self._path

    @alignment_path.setter
    def alignment_path(self, value):
        self._path = value

    def _dump(self):
        if self._dense is None:
            self._dense = self._store.dense_dump(self._gene_idx, self._free_params)
        return self._dense

    def _mat(self, k):
        return np.matrix(self._dump()["mats"][k])

    DP_M_matrix = property(lambda self: self._mat(0))
    DP_W_matrix = property(lambda self: self._mat(1))
    DP_V_matrix = property(lambda self: self._mat(2))
    DP_D_matrix = property(lambda self: self._mat(3))
    DP_I_matrix = property(lambda self: self._mat(4))

    @property
    def DP_util_matrix(self):
        """Object matrix whose interior cells hold ``[null, match_len, match_compression]``
        (OrgAlign.py:501); row/column 0 hold ``None``."""
        d = self._dump()
        n = self.T_len + 1
        out = np.empty((n, self.S_len + 1), dtype=object)
        util = d["util"]
        for i in range(1, n):
            for j in range(1, self.S_len + 1):
                out[i, j] = [np.asarray(util[0, i - 1, j - 1]), np.asarray(util[1, i - 1, j - 1]),
                             np.asarray(util[2, i - 1, j - 1])]
        return np.matrix(out)

    def _backpointers(self, k):
        d = self._dump()
        moves = {0: (-1, -1), 1: (0, -1), 2: (-1, 0), 3: (0, -1), 4: (-1, 0)}
        bp = d["bp"][k]
        out = []
        for i in range(bp.shape[0]):
            row = []
            for j in range(bp.shape[1]):
                if bp[i, j] < 0:
                    row.append([0, 0, 0] if (i == 0 and j == 0) else [np.inf, np.inf, np.inf])
                else:
                    row.append([i + moves[k][0], j + moves[k][1], int(bp[i, j])])
            out.append(row)
        return out

    backtrackers_M = property(lambda self: self._backpointers(0))
    backtrackers_W = property(lambda self: self._backpointers(1))
    backtrackers_V = property(lambda self: self._backpointers(2))
    backtrackers_D = property(lambda self: self._backpointers(3))
    backtrackers_I = property(lambda self: self._backpointers(4))

    def backtrack(self):
        return self.alignment_path

    def get_matched_regions(self):
        return legacy_matched_regions(self.alignment_str)


class AlignmentLandscape:
    """Per-gene landscape view (reference: OrgAlign.py:649-758)."""

    def __init__(self, fwd_DP, l_states, S_len, T_len):
        self._dp = fwd_DP
        self._ls = l_states  # uint8 [(T+1), (S+1)]
        self.S_len = S_len
        self.T_len = T_len
        self.the_5_state_machine = True
        self._str = None

    @property
    def FSA(self):
        return self._dp.FSA

    @property
    def alignment_path(self):
        return self._dp.alignment_path

    @property
    def L_matrix_states(self):
        """argmin state per DP cell as a ``<U1`` string matrix ('0'..'4'), what
        ``ClusterUtils.get_cluster_average_alignments`` compares against (OrgAlign.py:676-679, 747)."""
        if self._str is None:
            self._str = np.matrix(self._ls.astype("<U1"))
        return self._str

    @property
    def L_matrix(self):
        return np.matrix(self._dp._dump()["L"])

    def collate_fwd(self):
        return None

    def plot_alignment_landscape(self):
        import matplotlib.pyplot as plt
        import seaborn as sb
        fig, ax = plt.subplots(figsize=(5, 5))
        ax = sb.heatmap(self.L_matrix, square=True, cmap="jet", ax=ax)
        path = self.alignment_path
        ax.plot([p[1] + 0.5 for p in path], [p[0] + 0.5 for p in path], color="black", linewidth=3, alpha=0.5,
                linestyle="dashed")
        plt.xlabel("Reference", fontweight="bold")
        plt.ylabel("Query", fontweight="bold")
        plt.title("Alignment cost landscape")


# -- string post-processing shared by DP5 / AligmentObj ------------------------------------------------
_STEP = {"M": (1, 1), "W": (0, 1), "V": (1, 0), "D": (0, 1), "I": (1, 0)}


def path_from_string(s):
    """DP cells (i = query row, j = reference column) reached after each alignment character,
    listed from the end cell backwards like the reference's traceback."""
    i = j = 0
    cells = []
    for ch in s:
        di, dj = _STEP[ch]
        i += di
        j += dj
        cells.append([i, j])
    cells.reverse()
    return cells


def _runs(s, letter):
    return [(m.start(), m.end() - 1) for m in re.finditer(letter + "+", s)]


def legacy_matched_regions(s):
    """``DP5.get_matched_regions`` (OrgAlign.py:611-646): the three-state region walk the reference
    still runs on five-state strings.  W/V characters take no branch there, so the cursor then
    advances by the length of the *previously handled* run; region ids advance independently of
    the cursor.  Reproduced as observable behaviour (``match_regions_*`` of ``AligmentObj``)."""
    runs = {k: _runs(s, k) for k in "MID"}
    nxt = {"M": 0, "I": 0, "D": 0}
    i = j = c = 0
    step = None
    S_match, T_match, S_non, T_non = [], [], [], []
    while c < len(s):
        ch = s[c]
        if ch in nxt:
            a, b = runs[ch][nxt[ch]]
            nxt[ch] += 1
            step = b - a + 1
            if ch == "M":
                S_match.append([j, j + step - 1])
                T_match.append([i, i + step - 1])
                i += step
                j += step
            elif ch == "I":
                T_non.append([i, i + step - 1])
                i += step
            else:
                S_non.append([j, j + step - 1])
                j += step
        c += step
    return [S_match, T_match, S_non, T_non]
