"""Regenerates the golden fixtures from the reference checkout (/root/reference; build container only).

The reference is Java and cannot be executed here, so the fixtures are its test DATA files and facts derived from
its sample data, not outputs of its code:
  SimpleTest.fastq                  <- core/src/test/resources/fastq/SimpleTest.fastq        (T/fastq/FastqReaderTest.java:43-75)
  dengue1_test.out                  <- core/src/test/resources/projects/dengue1/test.out      (T/goals/refseq/DBGoalTest.java:127-142)
  taxtree_nodes.dmp / names.dmp     <- core/src/test/resources/taxtree/{nodes,names}.dmp      (T/match/FastqKMerMatcherTest.java:322-412)
  dengue1.fasta / dengue1_test.fastq <- core/src/test/resources/projects/dengue1/{dengue1.fasta,test.fastq} (same test)
  sample_fastq_read_lengths.txt.gz  <- read lengths of data/projects/human_virus/fastq/sample.fastq.gz (README.md:169 totals)
  human_virus_sample.fastq.gz       <- data/projects/human_virus/fastq/sample.fastq.gz itself (6565 real reads, variable length,
                                       real headers and qualities; README.md:169: 6565 reads / 658 255 bps / 461 305 k-mers)
"""
import gzip
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    res = os.path.join(REF, "core/src/test/resources")
    shutil.copy(os.path.join(res, "fastq/SimpleTest.fastq"), os.path.join(HERE, "SimpleTest.fastq"))
    shutil.copy(os.path.join(res, "projects/dengue1/test.out"), os.path.join(HERE, "dengue1_test.out"))
    shutil.copy(os.path.join(res, "taxtree/nodes.dmp"), os.path.join(HERE, "taxtree_nodes.dmp"))
    shutil.copy(os.path.join(res, "taxtree/names.dmp"), os.path.join(HERE, "taxtree_names.dmp"))
    shutil.copy(os.path.join(res, "projects/dengue1/dengue1.fasta"), os.path.join(HERE, "dengue1.fasta"))
    shutil.copy(os.path.join(res, "projects/dengue1/test.fastq"), os.path.join(HERE, "dengue1_test.fastq"))
    shutil.copy(os.path.join(REF, "data/projects/human_virus/fastq/sample.fastq.gz"), os.path.join(HERE, "human_virus_sample.fastq.gz"))
    lens = []
    with gzip.open(os.path.join(REF, "data/projects/human_virus/fastq/sample.fastq.gz"), "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(1, len(lines), 4):
        lens.append(len(lines[i]))
    with gzip.open(os.path.join(HERE, "sample_fastq_read_lengths.txt.gz"), "wb") as f:
        f.write(("\n".join(str(x) for x in lens) + "\n").encode())
    print("reads", len(lens), "bps", sum(lens))
