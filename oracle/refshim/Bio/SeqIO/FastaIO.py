def SimpleFastaParser(handle):
    title, chunks = None, []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield title, "".join(chunks)
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.strip())
    if title is not None:
        yield title, "".join(chunks)
