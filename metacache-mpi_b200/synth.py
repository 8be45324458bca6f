"""Seeded synthetic genomes, reads and taxonomy of the shapes BASELINE.json names
(SURVEY.md section 8d).  numpy only; deterministic per (seed, index)."""
import os

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, np.uint8)
COMP[:] = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[a] = b

# taxonomy.h:61-84
RANK = {"sequence": 0, "form": 1, "variety": 2, "subspecies": 3, "species": 4, "subgenus": 5, "genus": 6,
        "subtribe": 7, "tribe": 8, "subfamily": 9, "family": 10, "suborder": 11, "order": 12, "subclass": 13,
        "class": 14, "subphylum": 15, "phylum": 16, "subkingdom": 17, "kingdom": 18, "domain": 19, "root": 20,
        "none": 21}
RANK_NAMES = {v: k for k, v in RANK.items()}

GENOME_SEED = 20261018


def random_bases(rng, n):
    return BASES[rng.integers(0, 4, size=n, dtype=np.uint8)]


def genome(g, length, n_runs=True, lower=True):
    """genome g: even g = iid uniform ACGT; odd g = genome g-1 with 5 % substitutions;
    0.1 % of positions covered by runs of 50 'N'; 1 % of positions lower-cased"""
    base_g = g - (g % 2)
    rng = np.random.Generator(np.random.PCG64(GENOME_SEED + base_g))
    s = random_bases(rng, length)
    if g % 2 == 1:
        r2 = np.random.Generator(np.random.PCG64(GENOME_SEED + g))
        nsub = int(0.05 * length)
        pos = r2.integers(0, length, size=nsub)
        s[pos] = random_bases(r2, nsub)
    r3 = np.random.Generator(np.random.PCG64(GENOME_SEED * 3 + g))
    if n_runs and length >= 1000:
        for p in r3.integers(0, max(length - 50, 1), size=max(int(0.001 * length / 50), 1)):
            s[p:p + 50] = ord("N")
    if lower:
        pos = r3.integers(0, length, size=int(0.01 * length))
        s[pos] |= 0x20
    return s


def make_genomes(G, length):
    return [genome(g, length) for g in range(G)]


def revcomp_ascii(s):
    return COMP[s[::-1]]


def make_read_pairs(genomes, npairs, readlen=150, seed=7, sub=0.01, insert=(300, 500), start_index=0):
    """pairs: genome uniform, insert size U[insert], mate1 = first readlen bases,
    mate2 = reverse complement of the last readlen bases, `sub` substitutions
    (1 in 5 an 'N').  Returns (buf1, offs1, buf2, offs2, truth_genome)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = len(genomes)
    lens = np.array([len(g) for g in genomes], np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    cat = np.concatenate(genomes) if G > 1 else genomes[0]
    gsel = rng.integers(0, G, size=npairs)
    ins = np.minimum(rng.integers(insert[0], insert[1] + 1, size=npairs), lens[gsel])
    rl = int(min(readlen, ins.min())) if npairs else readlen
    assert rl == readlen, "genomes shorter than the read length"
    p = (rng.random(npairs) * (lens[gsel] - ins + 1)).astype(np.int64) + starts[gsel]
    ar = np.arange(readlen, dtype=np.int64)
    b1 = np.empty(npairs * readlen, np.uint8)
    b2 = np.empty(npairs * readlen, np.uint8)
    step = 1 << 18
    for lo in range(0, npairs, step):
        hi = min(lo + step, npairs)
        b1[lo * readlen:hi * readlen] = cat[(p[lo:hi, None] + ar[None, :]).ravel()]
        # mate 2: reverse complement of the fragment's last readlen bases
        e = p[lo:hi] + ins[lo:hi] - 1
        b2[lo * readlen:hi * readlen] = COMP[cat[(e[:, None] - ar[None, :]).ravel()]]
    for b in (b1, b2):
        nsub = int(sub * b.size)
        pos = rng.integers(0, max(b.size, 1), size=nsub)
        newb = random_bases(rng, nsub)
        newb[rng.integers(0, 5, size=nsub) == 0] = ord("N")
        b[pos] = newb
    offs = np.arange(npairs + 1, dtype=np.int64) * readlen
    return b1, offs, b2, offs.copy(), gsel


def make_long_reads(genomes, nreads, readlen=10000, seed=11, sub=0.10):
    rng = np.random.Generator(np.random.PCG64(seed))
    G = len(genomes)
    gsel = rng.integers(0, G, size=nreads)
    buf = np.empty(nreads * readlen, np.uint8)
    for i in range(nreads):
        gs = genomes[gsel[i]]
        p = int(rng.integers(0, len(gs) - readlen + 1))
        r = gs[p:p + readlen]
        if rng.integers(0, 2):
            r = revcomp_ascii(r)
        buf[i * readlen:(i + 1) * readlen] = r
    nsub = int(sub * buf.size)
    pos = rng.integers(0, buf.size, size=nsub)
    buf[pos] = random_bases(rng, nsub)
    offs = np.arange(nreads + 1, dtype=np.int64) * readlen
    return buf, offs, gsel


def make_taxonomy(G):
    """1 root -> 2 domain -> family 1000+g//4 -> unranked 5000+g//2 -> genus 10000+g//2
    -> species 100000+g.  Returns (ids, parents, ranks, species_of_genome)"""
    ids, par, rk = [1, 2], [1, 1], [RANK["root"], RANK["domain"]]
    for f in range((G + 3) // 4):
        ids.append(1000 + f); par.append(2); rk.append(RANK["family"])
    for j in range((G + 1) // 2):
        ids.append(5000 + j); par.append(1000 + j // 2); rk.append(RANK["none"])
        ids.append(10000 + j); par.append(5000 + j); rk.append(RANK["genus"])
    for g in range(G):
        ids.append(100000 + g); par.append(10000 + g // 2); rk.append(RANK["species"])
    species = np.array([100000 + g for g in range(G)], np.int64)
    return (np.array(ids, np.int64), np.array(par, np.int64), np.array(rk, np.uint8), species)


# ---- file writers (for the command-line drop-in tests) ----------------------
def write_fasta(path, header, seq, width=80):
    with open(path, "wb") as f:
        f.write(b">" + header.encode() + b"\n")
        n = len(seq)
        nl = (n + width - 1) // width
        out = np.full(n + nl, ord("\n"), np.uint8)
        idx = np.arange(n)
        out[idx + idx // width] = seq
        f.write(out.tobytes())
        if n % width != 0 or n == 0:
            pass


def write_genome_files(dirname, genomes, species):
    os.makedirs(dirname, exist_ok=True)
    paths = []
    for g, s in enumerate(genomes):
        p = os.path.join(dirname, "g%05d.fna" % g)
        write_fasta(p, "SYN%06d.1 taxid=%d synthetic genome %d" % (g, species[g], g), s)
        paths.append(p)
    return paths


def write_fastq_pair(path1, path2, b1, o1, b2, o2, truth=None, start=0):
    for path, buf, offs, mate in ((path1, b1, o1, 1), (path2, b2, o2, 2)):
        with open(path, "wb") as f:
            for i in range(len(offs) - 1):
                s = buf[offs[i]:offs[i + 1]].tobytes()
                g = int(truth[i]) if truth is not None else 0
                f.write(b"@read%d_g%d/%d\n" % (start + i, g, mate))
                f.write(s + b"\n+\n" + b"I" * len(s) + b"\n")


def write_taxonomy_files(dirname, ids, parents, ranks):
    os.makedirs(dirname, exist_ok=True)
    ncbi_name = {RANK["root"]: "no rank", RANK["domain"]: "superkingdom", RANK["none"]: "no rank"}
    with open(os.path.join(dirname, "nodes.dmp"), "w") as f:
        for i, p, r in zip(ids, parents, ranks):
            rn = ncbi_name.get(int(r), RANK_NAMES[int(r)])
            f.write("%d\t|\t%d\t|\t%s\t|\t\t|\n" % (i, p, rn))
    with open(os.path.join(dirname, "names.dmp"), "w") as f:
        for i in ids:
            f.write("%d\t|\ttaxon%d\t|\t\t|\tscientific name\t|\n" % (i, i))


# ---- counter-based generator ("bases iid uniform from a counter-based PRNG, e.g. splitmix64") -------
# Same shapes as genome() / make_read_pairs() above, but every base is a pure function of (stream, genome, position):
# any slice can be produced anywhere -- on the GPU with torch ops at configs[2] / configs[3] size (20 Gbp, 50 M
# pairs), on the CPU with the same code for the reference arm and the tests.  Bench data only; the parity tests keep
# the numpy generator above.
_M64 = (1 << 64) - 1


def _i64(x):
    """python int (mod 2^64) -> the same bit pattern as a signed 64-bit value (torch has no uint64 arithmetic)"""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def splitmix64_py(x):
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _lsr(z, k):
    """logical shift right of int64 tensors"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64_t(x):
    """splitmix64 on torch int64 tensors (two's complement wrap-around == arithmetic mod 2^64)"""
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _key(stream, g):
    return _i64(splitmix64_py((GENOME_SEED << 8) + stream) ^ ((g & 0xFFFFFF) << 40))


def hash_py(stream, g, pos):
    return splitmix64_py((splitmix64_py((GENOME_SEED << 8) + stream) ^ ((g & 0xFFFFFF) << 40) ^ pos) & _M64)


def genome_fast(g, length, device="cpu", out=None):
    """genome g as a torch uint8 tensor: even g = iid uniform ACGT, odd g = genome g-1 with 5 % substitutions,
    0.1 % of the aligned 50-base blocks are runs of 'N', 1 % of the positions lower-cased"""
    import torch
    pos = torch.arange(length, dtype=torch.int64, device=device)
    acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    base_g = g - (g % 2)
    code = _lsr(splitmix64_t(pos ^ _key(0, base_g)), 62)
    if g % 2 == 1:
        sub = _lsr(splitmix64_t(pos ^ _key(1, g)), 32) < int(0.05 * (1 << 32))
        code = torch.where(sub, _lsr(splitmix64_t(pos ^ _key(2, g)), 62), code)
    s = acgt[code]
    nrun = _lsr(splitmix64_t((pos // 50) ^ _key(3, g)), 32) < int(0.001 * (1 << 32))
    s = torch.where(nrun, torch.full_like(s, 78), s)
    low = _lsr(splitmix64_t(pos ^ _key(4, g)), 32) < int(0.01 * (1 << 32))
    s = torch.where(low, s | 0x20, s)
    if out is not None:
        out.copy_(s)
        return out
    return s


def read_pairs_fast(genomes, glen, npairs, first_pair=0, readlen=150, seed=7, sub=0.01, insert=(300, 500), chunk=1 << 17):
    """`npairs` read pairs (global pair indices first_pair ...) sampled from `genomes`: a 1-D torch uint8 tensor of
    len(genomes) // glen genomes of glen bases back to back, on any device.  Same shape as make_read_pairs():
    insert size U[insert], mate 1 = first readlen bases of the fragment, mate 2 = reverse complement of its last
    readlen bases, `sub` substitutions of which 1 in 5 an 'N'.  Returns (b1, b2, genome index per pair) on that device."""
    import torch
    dev = genomes.device
    G = genomes.numel() // glen
    comp = torch.arange(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[a] = b
    acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    b1 = torch.empty(npairs * readlen, dtype=torch.uint8, device=dev)
    b2 = torch.empty(npairs * readlen, dtype=torch.uint8, device=dev)
    gsel_all = torch.empty(npairs, dtype=torch.int64, device=dev)
    ar = torch.arange(readlen, dtype=torch.int64, device=dev)
    k = [_i64(splitmix64_py((seed << 16) + i)) for i in range(8)]
    for lo in range(0, npairs, chunk):
        hi = min(lo + chunk, npairs)
        idx = torch.arange(first_pair + lo, first_pair + hi, dtype=torch.int64, device=dev)
        gsel = _lsr(splitmix64_t(idx ^ k[0]), 1) % G
        ins = insert[0] + _lsr(splitmix64_t(idx ^ k[1]), 1) % (insert[1] - insert[0] + 1)
        p = _lsr(splitmix64_t(idx ^ k[2]), 1) % (glen - ins + 1) + gsel * glen
        m1 = genomes[(p[:, None] + ar[None, :]).reshape(-1)]
        e = p + ins - 1
        m2 = comp[genomes[(e[:, None] - ar[None, :]).reshape(-1)].long()]
        for mate, m in ((0, m1), (1, m2)):
            j = (idx[:, None] * readlen + ar[None, :]).reshape(-1)
            h = splitmix64_t(j ^ k[3 + mate])
            hit = _lsr(h, 32) < int(sub * (1 << 32))
            nb = acgt[_lsr(splitmix64_t(j ^ k[5 + mate]), 62)]
            nb = torch.where((_lsr(h, 3) & 0xFFFF) % 5 == 0, torch.full_like(nb, 78), nb)
            m.copy_(torch.where(hit, nb, m))
        b1[lo * readlen:hi * readlen] = m1
        b2[lo * readlen:hi * readlen] = m2
        gsel_all[lo:hi] = gsel
    return b1, b2, gsel_all


def long_reads_fast(genomes, glen, nreads, first_read=0, readlen=10000, seed=11, sub=0.10, chunk=1 << 11):
    """`nreads` single-end reads of `readlen` bases, random strand, `sub` substitutions (configs[4] shape)"""
    import torch
    dev = genomes.device
    G = genomes.numel() // glen
    comp = torch.arange(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[a] = b
    acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.empty(nreads * readlen, dtype=torch.uint8, device=dev)
    ar = torch.arange(readlen, dtype=torch.int64, device=dev)
    k = [_i64(splitmix64_py((seed << 16) + 100 + i)) for i in range(6)]
    for lo in range(0, nreads, chunk):
        hi = min(lo + chunk, nreads)
        idx = torch.arange(first_read + lo, first_read + hi, dtype=torch.int64, device=dev)
        gsel = _lsr(splitmix64_t(idx ^ k[0]), 1) % G
        p = _lsr(splitmix64_t(idx ^ k[1]), 1) % (glen - readlen + 1) + gsel * glen
        rev = (_lsr(splitmix64_t(idx ^ k[2]), 40) & 1) == 1
        fwd = genomes[(p[:, None] + ar[None, :]).reshape(-1)].view(hi - lo, readlen)
        bwd = comp[genomes[((p + readlen - 1)[:, None] - ar[None, :]).reshape(-1)].long()].view(hi - lo, readlen)
        m = torch.where(rev[:, None], bwd, fwd).reshape(-1)
        j = (idx[:, None] * readlen + ar[None, :]).reshape(-1)
        hit = _lsr(splitmix64_t(j ^ k[3]), 32) < int(sub * (1 << 32))
        nb = acgt[_lsr(splitmix64_t(j ^ k[4]), 62)]
        out[lo * readlen:hi * readlen] = torch.where(hit, nb, m)
    return out
