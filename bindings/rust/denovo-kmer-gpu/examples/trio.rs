//! The loop INTEGRATION.md §3 describes, as code: what the reference's `main.rs` does around
//! `kmer.rs` / `counter.rs`, with the GPU counter in their place.  The BAM / VCF / FASTA side is
//! the reference's own (rust-htslib) and is represented here by the `ReadSource` trait.
use denovo_kmer_gpu::{comm_unique_id, variant_kmers, Candidate, Counter, Thresholds, CALL_DENOVO};

/// What the reference's BAM layer yields: decoded reads of one sample, in batches.
pub trait ReadSource {
    /// Fills `seq` (ASCII here; `record.seq().encoded` with `four_bit = true` skips the decode),
    /// `qual`, `offsets` (offsets[0] = 0) with the next batch; false at end of file.
    fn next_batch(&mut self, seq: &mut Vec<u8>, qual: &mut Vec<u8>, offsets: &mut Vec<u64>) -> bool;
}

pub fn run_trio(device: i32, k: i32, min_baseq: i32, thr: Thresholds, cands: &[Candidate<'_>],
                samples: &mut [Box<dyn ReadSource>; 3], rank_world: Option<(i32, i32, [u8; 128])>)
                -> denovo_kmer_gpu::Result<Vec<bool>> {
    let mut kc = Counter::new(device, k)?;
    kc.bind_thread_near_gpu()?;
    if let Some((rank, world, id)) = rank_world {
        kc.comm_init(&id, rank, world)?; // every rank scans its own slice of the read batches
    }
    kc.build_table(&variant_kmers(cands, k, true)?)?;
    let mut packer = kc.batch_packer(1 << 21, 150 << 21, min_baseq)?;
    let (mut seq, mut qual, mut off) = (Vec::new(), Vec::new(), Vec::new());
    for (sample, src) in samples.iter_mut().enumerate() {
        let mut in_flight = 0;
        while src.next_batch(&mut seq, &mut qual, &mut off) {
            if in_flight == 2 {
                kc.sync()?; // both pinned buffers are in use: wait before repacking one
                in_flight = 0;
            }
            let batch = packer.next(&seq, false, Some(qual.as_slice()), &off)?;
            kc.submit(batch, sample as i32)?;
            in_flight += 1;
        }
        kc.sync()?;
    }
    kc.counts_allreduce()?; // no-op without a communicator
    let res = kc.finalise(&thr)?;
    Ok(res.calls.iter().map(|c| c & CALL_DENOVO != 0).collect())
}

fn main() {
    // rank 0 of a multi-GPU job would make the id and send it round:
    let _ = comm_unique_id();
    eprintln!("see run_trio(); the reference's CLI supplies the candidates and the three BAM readers");
}
