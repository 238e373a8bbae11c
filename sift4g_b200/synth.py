"""Synthetic protein queries / databases of the shapes BASELINE.json names.

Residues are i.i.d. from a 20-letter background plus a small fraction of the rare letters
(B, J, O, U, X, Z) so the 26-letter code path is exercised; database lengths are log-normal;
homologs of the queries (substitutions + geometric indels) are planted so that the E-value filter
keeps hits and the traceback stage has work.  Everything is seeded and numpy-only (host side);
the bench's device-side generator in bench.py follows the same recipe.
"""
import numpy as np

# background frequencies over A..Z codes (20 standard amino acids, Robinson & Robinson-like)
_AA = "ACDEFGHIKLMNPQRSTVWY"
_FREQ = np.array([0.078, 0.019, 0.054, 0.063, 0.039, 0.074, 0.022, 0.051, 0.057, 0.090,
                  0.022, 0.045, 0.052, 0.043, 0.051, 0.071, 0.058, 0.064, 0.013, 0.032])
_RARE = "BJOUXZ"


def letter_table(rare_fraction=0.001):
    p = np.zeros(26)
    for ch, f in zip(_AA, _FREQ):
        p[ord(ch) - 65] = f
    p = p / p.sum() * (1.0 - rare_fraction)
    for ch in _RARE:
        p[ord(ch) - 65] = rare_fraction / len(_RARE)
    return p


def random_codes(rng, n, rare_fraction=0.001):
    return rng.choice(26, size=int(n), p=letter_table(rare_fraction)).astype(np.uint8)


def mutate(rng, codes, identity=0.7, indel_rate=0.02, max_indel=6):
    """Return a mutated copy: substitutions at rate (1-identity), indels of geometric length."""
    out = []
    i, n = 0, len(codes)
    p = letter_table(0.0)
    while i < n:
        r = rng.random()
        if r < indel_rate / 2:            # deletion
            i += min(int(rng.geometric(0.5)), max_indel)
            continue
        if r < indel_rate:                # insertion
            out.extend(rng.choice(26, size=min(int(rng.geometric(0.5)), max_indel), p=p).tolist())
        c = int(codes[i])
        if rng.random() > identity:
            c = int(rng.choice(26, p=p))
        out.append(c)
        i += 1
    if not out:
        out = [int(codes[0])]
    return np.array(out, dtype=np.uint8)


def pack(seqs):
    """list of uint8 arrays -> (concatenated codes, int64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    codes = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return codes, off


def make_dataset(seed, n_queries, n_db, q_len=(100, 1000), homologs=(3, 10), db_len_mu=5.6,
                 db_len_sigma=0.6, db_len_clip=(30, 35000), identity=(0.3, 0.95), rare_fraction=0.001,
                 flank=(0, 60)):
    """-> (queries list, database list).  Planted homologs sit at random database positions."""
    rng = np.random.default_rng(seed)
    queries = [random_codes(rng, rng.integers(q_len[0], q_len[1] + 1), rare_fraction) for _ in range(n_queries)]
    lens = np.clip(np.exp(rng.normal(db_len_mu, db_len_sigma, size=n_db)), *db_len_clip).astype(np.int64)
    db = [random_codes(rng, l, rare_fraction) for l in lens]
    for q in queries:
        nh = int(rng.integers(homologs[0], homologs[1] + 1))
        for _ in range(nh):
            if n_db == 0:
                break
            lo, hi = 0, len(q)
            if rng.random() < 0.5 and len(q) > 60:       # partial (domain-level) homolog
                lo = int(rng.integers(0, len(q) // 2)); hi = int(rng.integers(lo + 30, len(q) + 1))
            core = mutate(rng, q[lo:hi], identity=rng.uniform(*identity))
            left = random_codes(rng, rng.integers(flank[0], flank[1] + 1), rare_fraction)
            right = random_codes(rng, rng.integers(flank[0], flank[1] + 1), rare_fraction)
            db[int(rng.integers(0, n_db))] = np.concatenate([left, core, right]).astype(np.uint8)
    return queries, db


def write_fasta(path, seqs, prefix="S", width=60):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%08d\n" % (prefix, i))
            txt = (np.asarray(s, dtype=np.uint8) + 65).tobytes().decode()
            for a in range(0, len(txt), width):
                f.write(txt[a:a + width] + "\n")
