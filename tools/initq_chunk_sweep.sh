for c in 37888 75776 151552 303104; do echo "chunk $c"; DIINN_INITQ_CHUNK=$c python tools/run_decode.py c3 fp16 5 3 1 2>&1 | tail -1; done
