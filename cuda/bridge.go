// Package cuda is the cgo bridge between the reference's Go packages (pianopir, graphann) and
// libpacmann_cuda.so.  REVIEW-ONLY SOURCE: no Go toolchain exists in the build image, so this file is never
// compiled here; the boundary it binds is exercised through ctypes and the C++ host mirror (INTEGRATION.md).
// Every wrapper passes slice memory for the duration of the call only and aborts on error, mirroring the
// reference's own log.Fatalf on unrecoverable conditions.  There is no CPU fallback.
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../pacmann_b200 -lpacmann_cuda -Wl,-rpath,${SRCDIR}/../pacmann_b200
#include <stdlib.h>
#include "pacmann_cuda.h"
*/
import "C"

import (
	"log"
	"unsafe"
)

func must(rc C.int, what string) {
	if rc != 0 {
		log.Fatalf("%s: libpacmann_cuda error %d: %s", what, int(rc), C.GoString(C.pm_last_error()))
	}
}

// DB is a device-resident copy of the flat rawDB (pianopir/pir.go:28-39).
type DB struct {
	h        *C.pm_db
	Rows     uint64
	EntryU64 uint64
}

func NewDB(rawDB []uint64, rows, entryU64 uint64, device int) *DB {
	var h *C.pm_db
	must(C.pm_db_create((*C.uint64_t)(unsafe.Pointer(&rawDB[0])), C.uint64_t(rows), C.uint64_t(entryU64), C.int(device), &h), "pm_db_create")
	return &DB{h: h, Rows: rows, EntryU64: entryU64}
}

func (d *DB) Close() { C.pm_db_destroy(d.h); d.h = nil }

// ExpandKey replaces expandKeyAsm / GetLongKey (pianopir/util.go:117,167-171).
func ExpandKey(key *[16]byte) []uint32 {
	rk := make([]uint32, 44)
	must(C.pm_expand_key((*C.uint8_t)(unsafe.Pointer(&key[0])), (*C.uint32_t)(unsafe.Pointer(&rk[0]))), "pm_expand_key")
	return rk
}

// HintJob mirrors struct pm_hint_job for one PianoPIR instance with the Initialization numbering.
type HintJob struct {
	Row0, Rows, ChunkSize, SetSize uint64
	LongKey                        []uint32 // 44 words
	HintBegin, Hints               uint64
	Primary, BackupGroup           uint64
	Parity                         []uint64 // [Hints][EntryU64], primary hints first, then backup hints
}

// HintGen replaces the loops of PianoPIRClient.Preprocessing / UpdatePreprocessing (pianopir/pir.go:267-352) for
// one or many sub-PIRs (SimpleBatchPianoPIR.Preprocessing, batch-pir.go:119-155) in a single call.
func (d *DB) HintGen(jobs []HintJob) {
	cj := (*[1 << 20]C.pm_hint_job)(C.calloc(C.size_t(len(jobs)), C.size_t(unsafe.Sizeof(C.pm_hint_job{}))))
	defer C.free(unsafe.Pointer(cj))
	for i, j := range jobs {
		cj[i].row0, cj[i].n_rows = C.uint64_t(j.Row0), C.uint64_t(j.Rows)
		cj[i].chunk_size, cj[i].set_size = C.uint64_t(j.ChunkSize), C.uint64_t(j.SetSize)
		for k := 0; k < 44; k++ {
			cj[i].rk[k] = C.uint32_t(j.LongKey[k])
		}
		cj[i].hint_begin, cj[i].n_hints = C.uint64_t(j.HintBegin), C.uint64_t(j.Hints)
		cj[i].n_primary, cj[i].backup_group = C.uint64_t(j.Primary), C.uint64_t(j.BackupGroup)
		// NOTE: storing a Go pointer in C memory for the duration of one call needs runtime.Pinner (Go >= 1.21);
		// with older toolchains pass C.malloc'ed staging buffers and copy out.
		cj[i].parity_out = (*C.uint64_t)(unsafe.Pointer(&j.Parity[0]))
	}
	must(C.pm_hintgen(d.h, &cj[0], C.uint64_t(len(jobs))), "pm_hintgen")
}

// GatherRows fetches replacement values (pianopir/pir.go:345-349); idx >= rows yields the zero padding entry.
func (d *DB) GatherRows(row0, rows uint64, idx []uint64, out []uint64) {
	must(C.pm_gather_rows(d.h, C.uint64_t(row0), C.uint64_t(rows), (*C.uint64_t)(unsafe.Pointer(&idx[0])), C.uint64_t(len(idx)),
		(*C.uint64_t)(unsafe.Pointer(&out[0]))), "pm_gather_rows")
}

// AnswerBatch replaces q calls of PianoPIRServer.PrivateQuery (pianopir/pir.go:65-88) with one launch.
func (d *DB) AnswerBatch(row0, rows []uint64, chunk, set []uint32, offsets []uint32, stride uint64, out []uint64) {
	q := len(row0)
	must(C.pm_answer_batch(d.h, (*C.uint64_t)(unsafe.Pointer(&row0[0])), (*C.uint64_t)(unsafe.Pointer(&rows[0])),
		(*C.uint32_t)(unsafe.Pointer(&chunk[0])), (*C.uint32_t)(unsafe.Pointer(&set[0])),
		(*C.uint32_t)(unsafe.Pointer(&offsets[0])), C.uint64_t(stride), C.uint64_t(q),
		(*C.uint64_t)(unsafe.Pointer(&out[0]))), "pm_answer_batch")
}

// L2Query replaces a loop of graphann.L2Dist(v, query) over the vertices of one search step
// (graphann/search.go:132,204,215); bit-identical to L2DistanceSIMD's evaluation order.
func L2Query(vecs []float32, n, dim uint64, query []float32, out []float32, device int) {
	must(C.pm_l2_query((*C.float)(unsafe.Pointer(&vecs[0])), C.uint64_t(n), C.uint64_t(dim), (*C.float)(unsafe.Pointer(&query[0])),
		(*C.float)(unsafe.Pointer(&out[0])), C.int(device)), "pm_l2_query")
}

// L2Batch: distances from queries[q] to the vectors stored at the head of rows ids[q][k] of the resident table.
func (d *DB) L2Batch(dim uint64, queries []float32, nq uint64, ids []int64, k uint64, out []float32) {
	must(C.pm_l2_batch(d.h, C.uint64_t(dim), (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq),
		(*C.int64_t)(unsafe.Pointer(&ids[0])), C.uint64_t(k), (*C.float)(unsafe.Pointer(&out[0]))), "pm_l2_batch")
}

// IPScan replaces the InnerProduct scan loop of TestInnerProduct (graphann/graphann_test.go:268-273).
func (d *DB) IPScan(dim uint64, queries []uint32, nq uint64, checksum []uint32) {
	must(C.pm_ip_u32_scan(d.h, C.uint64_t(dim), (*C.uint32_t)(unsafe.Pointer(&queries[0])), C.uint64_t(nq),
		(*C.uint32_t)(unsafe.Pointer(&checksum[0])), nil), "pm_ip_u32_scan")
}

// L2IDPairs: out[p] = L2Dist(vector of row a[p], vector of row b[p]) over the resident table -- every L2Dist call of
// robustPrune (graphann/build_graph.go:169-236) for a batch of vertices in one launch.
func (d *DB) L2IDPairs(dim uint64, a, b []int64, out []float32) {
	must(C.pm_l2_idpairs(d.h, C.uint64_t(dim), (*C.int64_t)(unsafe.Pointer(&a[0])), (*C.int64_t)(unsafe.Pointer(&b[0])),
		C.uint64_t(len(a)), (*C.float)(unsafe.Pointer(&out[0]))), "pm_l2_idpairs")
}

// Client is the GPU-resident form of the PianoPIR client(s) of one SimpleBatchPianoPIR -- or of several of them
// (lanes x PartitionNum parts) when many users are served in lock step (INTEGRATION.md, steps 7 and 8).
type Client struct{ h *C.pm_client }

type ClientPart = C.pm_client_part
type ClientQuery = C.pm_client_query

func (d *DB) NewClient(parts []ClientPart) *Client {
	var h *C.pm_client
	must(C.pm_client_create(d.h, &parts[0], C.uint64_t(len(parts)), &h), "pm_client_create")
	return &Client{h}
}
func (c *Client) Close() { C.pm_client_destroy(c.h); c.h = nil }

// Preprocess = Initialization + Preprocessing of the listed parts on the device (pir.go:203-255, 267-352).
func (c *Client) Preprocess(partIDs []uint32, longKeys []uint32, replSeeds []uint64, skipPrep bool) {
	sp := C.int(0)
	if skipPrep {
		sp = 1
	}
	must(C.pm_client_preprocess(c.h, (*C.uint32_t)(unsafe.Pointer(&partIDs[0])), C.uint64_t(len(partIDs)),
		(*C.uint32_t)(unsafe.Pointer(&longKeys[0])), (*C.uint64_t)(unsafe.Pointer(&replSeeds[0])), sp), "pm_client_preprocess")
}

// QueryBatchL2M answers the sub-queries of one search step of all lanes: entries, status codes and the distance of
// every answered vector to its own lane's query vector (queryVecs[vecID[i]]).
func (c *Client) QueryBatchL2M(q []ClientQuery, out []uint64, status []int32, queryVecs []float32, nVecs uint64, vecID []uint32,
	dim uint64, dist []float32) {
	must(C.pm_client_query_batch_l2m(c.h, &q[0], C.uint64_t(len(q)), (*C.uint64_t)(unsafe.Pointer(&out[0])),
		(*C.int32_t)(unsafe.Pointer(&status[0])), (*C.float)(unsafe.Pointer(&queryVecs[0])), C.uint64_t(nVecs),
		(*C.uint32_t)(unsafe.Pointer(&vecID[0])), C.uint64_t(dim), (*C.float)(unsafe.Pointer(&dist[0]))), "pm_client_query_batch_l2m")
}

// Search is the GPU-resident lock-step SearchKNN of the clients ("lanes") of one Client (INTEGRATION.md step 9): the
// frontier of graphann.SearchKNN (search.go:114-234) and SimpleBatchPianoPIR.Query's bookkeeping live in HBM; a round is
// Begin, maxStep x Fetch (enqueue only), Finish.  The Go side keeps what needs no entry data: the batch budget of every
// lane (batch-pir.go:239-245: call Apply, then Client.Preprocess for the due lanes, then go on with Fetch(false)).
type Search struct{ h *C.pm_search }

type SearchConfig = C.pm_search_config

func (c *Client) NewSearch(cfg *SearchConfig) *Search {
	var h *C.pm_search
	must(C.pm_search_create(c.h, cfg, &h), "pm_search_create")
	return &Search{h}
}
func (s *Search) Close() { C.pm_search_destroy(s.h); s.h = nil }

func (s *Search) SetStart(lane uint32, ids []int64, vectors []float32, neighbors []int32) {
	must(C.pm_search_set_start(s.h, C.uint32_t(lane), (*C.int64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&vectors[0])),
		(*C.int32_t)(unsafe.Pointer(&neighbors[0]))), "pm_search_set_start")
}
func (s *Search) SetDummySeed(lane uint32, seeds []uint64) {
	must(C.pm_search_set_dummy_seed(s.h, C.uint32_t(lane), (*C.uint64_t)(unsafe.Pointer(&seeds[0]))), "pm_search_set_dummy_seed")
}
func (s *Search) Begin(lanes []uint32, queries []float32, randSeeds []uint64, k uint64, benchmarking bool) {
	b := C.int(0)
	if benchmarking {
		b = 1
	}
	must(C.pm_search_begin(s.h, (*C.uint32_t)(unsafe.Pointer(&lanes[0])), C.uint64_t(len(lanes)), (*C.float)(unsafe.Pointer(&queries[0])),
		(*C.uint64_t)(unsafe.Pointer(&randSeeds[0])), C.uint64_t(k), b), "pm_search_begin")
}
func (s *Search) Fetch(applyPrevious bool) {
	a := C.int(0)
	if applyPrevious {
		a = 1
	}
	must(C.pm_search_fetch(s.h, a), "pm_search_fetch")
}
func (s *Search) Apply() { must(C.pm_search_apply(s.h), "pm_search_apply") }
func (s *Search) Finish(applyPrevious bool, ret, stepRet []int64, stats, finished []uint64) {
	a := C.int(0)
	if applyPrevious {
		a = 1
	}
	must(C.pm_search_finish(s.h, a, (*C.int64_t)(unsafe.Pointer(&ret[0])), (*C.int64_t)(unsafe.Pointer(&stepRet[0])),
		(*C.uint64_t)(unsafe.Pointer(&stats[0])), (*C.uint64_t)(unsafe.Pointer(&finished[0]))), "pm_search_finish")
}
