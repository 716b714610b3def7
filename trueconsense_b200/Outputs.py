"""Output writers with the reference's API (TrueConsense/Outputs.py): consensus FASTA, the
VCF-like variant list and the corrected GFF.

Pure host formatting around the GPU path: ``WriteOutputs`` needs the consensus with and without
insertions plus the insertion list (Outputs.py:94-105 runs BuildConsensus twice and ListInserts a
third time); here the count table is ranked once per (mincov, ambiguity) setting by the call kernel
and the insertion columns are piled up once per BAM handle (Events.ListInserts caches them), so
the three passes cost one.  Every quirk of the VCF writer is kept:
records are produced from the same comparisons in the same order.
"""
from __future__ import annotations

import sys
from datetime import date

from .Coverage import GetCoverage
from .Events import ListInserts
from .indexing import Readbam
from .Sequences import BuildConsensus

GFF_COLUMNS = ["seqid", "source", "type", "start", "end", "score", "strand", "phase", "attributes"]


def _first_fasta_record(path):
    """(id, list of residues) of the first record — what Outputs.py:107-113 keeps of SeqIO.parse."""
    rec_id, chunks, seen = None, [], False
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if seen:
                    break
                seen = True
                words = line[1:].split(None, 1)
                rec_id = words[0] if words else ""
            elif seen:
                chunks.append("".join(line.split()))
    if not seen:
        raise UnboundLocalError("cannot access local variable 'reflist' where it is not associated with a value")
    return rec_id, list("".join(chunks))


def _gff_line(feature: dict) -> str:
    """One output line for a feature dict (Outputs.py:31-57): the eight fixed columns keep their
    value, everything else — including the parsed ``attributes`` string — is folded into the
    attributes column as key=value pairs (later keys overwrite earlier ones)."""
    fixed = GFF_COLUMNS[:-1]
    cleaned: dict[str, str] = {}
    extra: dict[str, str] = {}
    for key, value in feature.items():
        low = str(key).lower()
        if low in fixed:
            cleaned[low] = str(value)
        else:
            extra[low] = str(value)
    merged: dict[str, str] = {}
    for key, value in extra.items():
        if key != "attributes":
            merged[key] = value
            continue
        for item in value.split(";"):
            if item == "":
                continue
            k, v = item.split("=")           # ValueError for 'a=b=c' or a bare word, like the reference
            merged[k] = v
    cleaned["attributes"] = ";".join(f"{k}={v}" for k, v in merged.items())
    assert list(cleaned.keys()) == GFF_COLUMNS
    return "\t".join(cleaned.values()) + "\n"


def WriteGFF(gffheader, gffdict, output_gff, name):
    """Outputs.py:13-71."""
    with open(output_gff, "w") as out:
        out.write(gffheader.raw_text)
        for feature in gffdict.values():
            out.write(_gff_line(feature))


def WriteOutputs(mincov, iDict, uGffDict, inputbam, IncludeAmbig, output_vcf, name, ref, output_gff, gffheader,
                 output_consensus):
    """Outputs.py:74-183."""
    today = date.today().strftime("%Y%m%d")
    bam = Readbam(inputbam)
    consensus, newgff = BuildConsensus(mincov, iDict, uGffDict, IncludeAmbig, bam, True)
    consensus_noinsert = BuildConsensus(mincov, iDict, uGffDict, IncludeAmbig, bam, False)[0]

    if output_gff is not None:
        WriteGFF(gffheader, newgff, output_gff, name)

    if output_vcf is not None:
        inserts = ListInserts(iDict, mincov, bam)
        refID, reflist = _first_fasta_record(ref)
        seqlist = list(consensus_noinsert.upper())
        with open(output_vcf, "w") as out:
            out.write("##fileformat=VCFv4.3\n"
                      f"##fileDate={today}\n"
                      f"##source='TrueConsense {' '.join(sys.argv[1:])}'\n"
                      f"##reference='{ref}'\n"
                      f"##contig=<ID={refID}>\n"
                      '##INFO=<ID=DP,Number=1,Type=Integer,Description="Read Depth">\n'
                      '##INFO=<ID=INDEL,Number=0,Type=Flag,Description="Indicates that the variant is an INDEL.">\n'
                      "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
            # records are streamed, so a failure part-way leaves the same partial file the reference leaves
            hasinserts, insertpositions = inserts
            in_deletion = set()
            for i in range(len(reflist)):
                if i in in_deletion:
                    continue
                if reflist[i] != seqlist[i]:
                    if seqlist[i] == "-":
                        b = i
                        gap = []
                        while seqlist[b] == "-":
                            gap.append(reflist[b])
                            in_deletion.add(b)
                            b += 1
                        refallele = str(reflist[i - 1] + "".join(gap))
                        depth = GetCoverage(iDict, i + 1)
                        out.write(f"{refID}\t{i}\t.\t{refallele}\t{seqlist[i - 1]}\t.\tPASS\tDP={depth};INDEL\n")
                    else:
                        p = 1 if i < 2 else i
                        depth = GetCoverage(iDict, p + 1)
                        out.write(f"{refID}\t{i + 1}\t.\t{reflist[i]}\t{seqlist[i]}\t.\tPASS\tDP={depth}\n")
                if hasinserts is True:
                    for lposition in insertpositions:
                        if i != lposition:
                            continue
                        depth = GetCoverage(iDict, i + 1)
                        if depth > mincov:
                            for size in insertpositions.get(lposition):
                                alt = seqlist[i] + str(insertpositions.get(lposition).get(size))
                                out.write(f"{refID}\t{i}\t.\t{reflist[i]}\t{alt}\t.\tPASS\tDP={depth};INDEL\n")

    with open(output_consensus, "w") as out:
        out.write(f">{name} mincov={mincov}\n{consensus}\n")
