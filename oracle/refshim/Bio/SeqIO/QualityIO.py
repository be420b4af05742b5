def FastqGeneralIterator(handle):
    while True:
        head = handle.readline()
        if not head:
            return
        if not head.strip():
            continue
        if head[0] != "@":
            raise ValueError("Records in Fastq files should start with '@' character")
        seq = handle.readline().rstrip("\r\n")
        plus = handle.readline()
        if not plus or plus[0] != "+":
            raise ValueError("Missing '+' line in FASTQ record")
        qual = handle.readline().rstrip("\r\n")
        if len(seq) != len(qual):
            raise ValueError("Lengths of sequence and quality values differs")
        yield head[1:].rstrip(), seq, qual
