#!/bin/bash
timeout 600 python scripts/tok_chunk_sweep.py 56832 75776 113664 151552 294912 2>&1 | tail -n 8
