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
