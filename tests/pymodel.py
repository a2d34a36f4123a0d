"""Independent pure-Python (string-level) model of the path, used to cross-check the C oracle's
bit-level code on tiny inputs.  Written from the rules in SURVEY.md Appendix B, not from mh_oracle.c.
Everything works on strings over ACGT plus '$'; ordering '$' < 'A' < 'C' < 'G' < 'T' for the k-th
character, and b = '$' sorts AFTER real b (megahit stores $ as 4 in the b field)."""
from collections import Counter

COMP = str.maketrans("ACGT", "TGCA")
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def revcomp(s):
    return s.translate(COMP)[::-1]


def trim_n(s):
    """megahit FastxReader::TrimN: first N-free segment."""
    b = None
    i = 0
    for i, ch in enumerate(s):
        if ch in "Nn":
            if b is not None:
                break
        elif b is None:
            b = i
    else:
        i = len(s)
    if b is None:
        return ""
    return s[b:i]


def clean(s):
    return "".join(ch if ch in "ACGT" else ("G" if ch in "Nn" else "A") for ch in s.upper().replace("n", "N"))


def count_edges(reads, k, m):
    """reads: true-orientation ACGT strings. Returns sorted list of (stored canonical (k+1)-mer, count)."""
    c = Counter()
    for r in reads:
        s = r[::-1]
        for i in range(len(s) - k):
            e = s[i:i + k + 1]
            c[min(e, revcomp(e))] += 1
    return sorted((e, min(n, 65535)) for e, n in c.items() if n >= m)


def pack(s, nwords):
    v = 0
    for ch in s:
        v = (v << 2) | CODE[ch]
    v <<= 32 * nwords - 2 * len(s)
    return [(v >> (32 * (nwords - 1 - i))) & 0xFFFFFFFF for i in range(nwords)]


def edge_words(e, cnt, k):
    nw = (2 * (k + 1) + 16 + 31) // 32
    w = pack(e, nw)
    w[-1] |= cnt
    return w


def _item_sort_key(kmer, a_is_dollar, b, inv_mul):
    # chars (with $ position as 'A'=0), then non-dollar flag, then b code (A..T = 0..3, $ = 4), then inverted mult
    chars = kmer.replace("$", "A")
    return ([CODE[c] for c in chars], 0 if a_is_dollar else 1, 4 if b == "$" else CODE[b], inv_mul)


def sdbg_from_seqs(seqs, k, mode="seq2sdbg"):
    """seqs: list of (stored-orientation string, mult). Returns list of dict(label, w, last, tip, mul)."""
    items = []
    for s, mult in seqs:
        if len(s) < k + 1:
            continue
        for t in (s, revcomp(s)):
            L = len(t)
            for o in range(0, L - k + 2):
                full = o + k <= L
                kmer = t[o:o + k] if full else t[o:o + k - 1] + "$"
                b = "$" if o == 0 else t[o - 1]
                cnt = mult if (o > 0 and full) else 0
                items.append((kmer, b, cnt))
    return _emit(items, k, mode)


def _emit(items, k, mode):
    items = sorted(items, key=lambda it: _item_sort_key(it[0], it[0].endswith("$"), it[1], 65535 - it[2]))
    out = []
    i = 0
    n = len(items)
    while i < n:
        j = i
        while j < n and items[j][0][:k - 1] == items[i][0][:k - 1]:
            j += 1
        group = items[i:j]
        solid_a = {it[0][-1] for it in group if it[0][-1] != "$" and it[1] != "$"}
        solid_b = {it[1] for it in group if it[0][-1] != "$" and it[1] != "$"}
        # merge identical (a,b) keeping the first (largest multiplicity) / counting duplicates
        merged = []
        for it in group:
            if merged and merged[-1][0] == it[0] and merged[-1][1] == it[1]:
                merged[-1][3] += 1
            else:
                merged.append([it[0], it[1], it[2], 1])
        kept = []
        for kmer, b, cnt, dup in merged:
            a = kmer[-1]
            if a == "$" and b in solid_b:
                continue
            if b == "$" and a in solid_a:
                continue
            kept.append((kmer, b, cnt, dup))
        seen_b = set()
        for idx, (kmer, b, cnt, dup) in enumerate(kept):
            a = kmer[-1]
            w = 0 if b == "$" else (CODE[b] + 5 if b in seen_b else CODE[b] + 1)
            seen_b.add(b)
            last = 0 if a == "$" else int(all(x[0][-1] != a for x in kept[idx + 1:]))
            tip = int(a == "$")
            if mode == "seq2sdbg":
                mul = cnt
            else:
                mul = 0 if (a == "$" or b == "$") else min(dup, 65535)
            out.append(dict(label=kmer, w=w, last=last, tip=tip, mul=mul, b=b))
        i = j
    return out


def sdbg_from_reads(reads, k, m):
    """read2sdbg stage-2 style: items only from solid runs, palindromes once, multiplicity by duplicates."""
    solid = {e for e, _ in count_edges(reads, k, m)} if m > 1 else None
    items = []
    for r in reads:
        s = r[::-1]
        npos = len(s) - k
        if npos <= 0:
            continue
        sol = [solid is None or min(s[i:i + k + 1], revcomp(s[i:i + k + 1])) in solid for i in range(npos)]
        for i in range(npos):
            if not sol[i]:
                continue
            e = s[i:i + k + 1]
            rc = revcomp(e)
            strands = [e] if e == rc else [e, rc]
            first = i == 0 or not sol[i - 1]
            lastp = i == npos - 1 or not sol[i + 1]
            for si, t in enumerate(strands):
                items.append((t[1:], t[0], 0))
                head = first if si == 0 else lastp
                tail = lastp if si == 0 else first
                if head:
                    items.append((t[:k], "$", 0))
                if tail:
                    items.append((t[2:] + "$", t[1], 0))
    return _emit(items, k, "read2sdbg")
