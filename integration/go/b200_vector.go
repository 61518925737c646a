//go:build cgo

// Package provider — B200 vector store: a VectorStore provider
// (provider/online.go:55-64) whose Nearest() is an in-process cgo call into
// libehb200.so instead of a network hop to Redis (provider/redis.go:454-493) or
// Pinecone (provider/pinecone.go:348-373).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Go toolchain here
// (`go version` -> not found).  The call pattern it produces — many OS threads, one query per call — is
// exercised against the same C ABI by tests/cpp/concurrent_search.c (64 pthreads).  The file is written against the interfaces
// cited inline and is meant to be dropped into the reference at
// provider/b200_vector.go (see INTEGRATION.md for the registration hunks).
package provider

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../embeddinghub_b200 -lehb200
#include <stdlib.h>
#include "ehb200.h"
*/
import "C"

import (
	"encoding/json"
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/featureform/fferr"
	pc "github.com/featureform/provider/provider_config"
	pt "github.com/featureform/provider/provider_type"
	"github.com/featureform/provider/types"
)

// B200VectorConfig mirrors the JSON config style of provider_config/pinecone_config.go:19-43.
type B200VectorConfig struct {
	Metric         string `json:"Metric"` // "l2" (embeddingstore default), "ip", "cosine" (what redis.go:253 / pinecone.go:252 use)
	M              uint32 `json:"M"`
	EfConstruction uint32 `json:"EfConstruction"`
	EfSearch       uint32 `json:"EfSearch"`
	Device         int32  `json:"Device"`
	Devices        []int32 `json:"Devices"` // more than one entry: range-sharded over those GPUs (ehb_sharded_*)
	Seed           uint64 `json:"Seed"`
}

// Serialize — the counterpart every provider config has (provider_config/pinecone_config.go:29-35).
func (c *B200VectorConfig) Serialize() pc.SerializedConfig {
	config, err := json.Marshal(c)
	if err != nil {
		panic(err)
	}
	return config
}

func (c *B200VectorConfig) Deserialize(config pc.SerializedConfig) error {
	if err := json.Unmarshal(config, c); err != nil {
		return fferr.NewInternalError(err)
	}
	return nil
}

type b200VectorStore struct {
	cfg    B200VectorConfig
	mu     sync.RWMutex
	tables map[string]*b200Table
	BaseProvider
}

// b200VectorStoreFactory is registered with RegisterFactory(pt.B200VectorOnline, ...) (provider/provider.go:90-100).
func b200VectorStoreFactory(serialized pc.SerializedConfig) (Provider, error) {
	cfg := B200VectorConfig{Metric: "cosine", M: 16, EfConstruction: 200, EfSearch: 64, Seed: 100}
	if err := cfg.Deserialize(serialized); err != nil {
		return nil, err
	}
	return &b200VectorStore{cfg: cfg, tables: map[string]*b200Table{},
		BaseProvider: BaseProvider{ProviderType: pt.B200VectorOnline, ProviderConfig: serialized}}, nil
}

func (s *b200VectorStore) AsOnlineStore() (OnlineStore, error) { return s, nil }
func (s *b200VectorStore) Close() error {
	s.mu.Lock()
	defer s.mu.Unlock()
	for k, t := range s.tables {
		C.ehb_index_destroy(t.ix)
		delete(s.tables, k)
	}
	return nil
}

func tableKey(feature, variant string) string { return feature + "\x00" + variant }

// ehb_last_error() is thread-local in the library and goroutines migrate between OS threads, so every call
// whose failure message is read runs between runtime.LockOSThread / UnlockOSThread (see callLocked).
func lastError(rc C.int, what string) error {
	return fferr.NewInternalError(fmt.Errorf("ehb200 %s: status %d: %s", what, int(rc), C.GoString(C.ehb_last_error())))
}

// callLocked runs fn (one cgo call returning a status) and, on failure, reads the thread-local error text
// on the same OS thread.
func callLocked(what string, fn func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := fn(); rc != C.EHB_OK {
		return lastError(rc, what)
	}
	return nil
}

// CreateIndex — provider/online.go:56.  dims come from types.VectorType (provider/types/value_type.go:96-100).
func (s *b200VectorStore) CreateIndex(feature, variant string, vectorType types.VectorType) (VectorStoreTable, error) {
	s.mu.Lock()
	defer s.mu.Unlock()
	key := tableKey(feature, variant)
	if _, ok := s.tables[key]; ok {
		return nil, fferr.NewDatasetAlreadyExistsError(feature, variant, nil)
	}
	var p C.ehb_params
	C.ehb_params_default(&p, C.uint32_t(vectorType.Dimension))
	switch s.cfg.Metric {
	case "ip":
		p.metric = C.EHB_IP
	case "l2":
		p.metric = C.EHB_L2
	default:
		p.metric = C.EHB_COSINE
	}
	p.M, p.ef_construction, p.ef_search = C.uint32_t(s.cfg.M), C.uint32_t(s.cfg.EfConstruction), C.uint32_t(s.cfg.EfSearch)
	p.seed, p.device = C.uint64_t(s.cfg.Seed), C.int32_t(s.cfg.Device)
	var ix *C.ehb_index
	if err := callLocked("create", func() C.int { return C.ehb_index_create(&p, &ix) }); err != nil {
		return nil, err
	}
	t := &b200Table{ix: ix, dim: int(vectorType.Dimension), labels: map[string]uint64{},
		cosine: p.metric == C.EHB_COSINE, originals: map[string][]float32{}}
	s.tables[key] = t
	return t, nil
}

func (s *b200VectorStore) DeleteIndex(feature, variant string) error { return s.DeleteTable(feature, variant) }

// OnlineStore — provider/online.go:42-48
func (s *b200VectorStore) GetTable(feature, variant string) (OnlineStoreTable, error) {
	s.mu.RLock()
	defer s.mu.RUnlock()
	t, ok := s.tables[tableKey(feature, variant)]
	if !ok {
		return nil, fferr.NewDatasetNotFoundError(feature, variant, nil)
	}
	return t, nil
}
func (s *b200VectorStore) CreateTable(feature, variant string, valueType types.ValueType) (OnlineStoreTable, error) {
	vt, ok := valueType.(types.VectorType)
	if !ok {
		return nil, fferr.NewInvalidArgumentError(fmt.Errorf("b200 vector store holds vectors only, got %T", valueType))
	}
	if t, err := s.GetTable(feature, variant); err == nil { // vectorstore_test.go:133-141 calls CreateIndex then CreateTable
		return t, nil
	}
	return s.CreateIndex(feature, variant, vt)
}
func (s *b200VectorStore) DeleteTable(feature, variant string) error {
	s.mu.Lock()
	defer s.mu.Unlock()
	key := tableKey(feature, variant)
	t, ok := s.tables[key]
	if !ok {
		return fferr.NewDatasetNotFoundError(feature, variant, nil)
	}
	C.ehb_index_destroy(t.ix)
	delete(s.tables, key)
	return nil
}

type b200Table struct {
	ix     *C.ehb_index
	dim    int
	mu     sync.RWMutex
	labels map[string]uint64
	keys   []string
	// The index stores cosine rows normalised (hnswlib's convention).  Get must return what Set stored
	// (vectorstore_test.go testGetSet: reflect.DeepEqual; Redis and Pinecone return the original), so for
	// cosine tables the original rows are kept host-side, like offlinehub keeps its own copy.
	cosine    bool
	originals map[string][]float32
}

// Set — provider/online.go:51; value must be []float32 (cf. pinecone.go:199-207, redis.go:407-413).
func (t *b200Table) Set(entity string, value interface{}) error {
	vec, ok := value.([]float32)
	if !ok || len(vec) != t.dim {
		return fferr.NewInvalidArgumentError(fmt.Errorf("expected []float32 of length %d, got %T", t.dim, value))
	}
	t.mu.Lock()
	label, seen := t.labels[entity]
	if !seen {
		label = uint64(len(t.keys))
		t.labels[entity] = label
		t.keys = append(t.keys, entity)
	}
	if t.cosine {
		t.originals[entity] = append([]float32(nil), vec...)
	}
	t.mu.Unlock()
	// Go memory is only borrowed for the duration of the call (cgo pointer rules): the library copies.
	return callLocked("add", func() C.int {
		return C.ehb_index_add(t.ix, 1, (*C.float)(unsafe.Pointer(&vec[0])), (*C.uint64_t)(unsafe.Pointer(&label)))
	})
}

// Delete — tombstones the entity (docs: space.delete / multidelete; ehb_index_remove).
func (t *b200Table) Delete(entity string) error {
	t.mu.Lock()
	label, ok := t.labels[entity]
	delete(t.originals, entity)
	t.mu.Unlock()
	if !ok {
		return fferr.NewEntityNotFoundError("", "", entity, nil)
	}
	return callLocked("remove", func() C.int { return C.ehb_index_remove(t.ix, 1, (*C.uint64_t)(unsafe.Pointer(&label))) })
}

// Get — provider/online.go:52
func (t *b200Table) Get(entity string) (interface{}, error) {
	t.mu.RLock()
	label, ok := t.labels[entity]
	var orig []float32
	if ok && t.cosine {
		orig = t.originals[entity]
	}
	t.mu.RUnlock()
	if !ok {
		return nil, fferr.NewEntityNotFoundError("", "", entity, nil)
	}
	if orig != nil {
		return append([]float32(nil), orig...), nil
	}
	out := make([]float32, t.dim)
	if err := callLocked("get", func() C.int {
		return C.ehb_index_get(t.ix, C.uint64_t(label), (*C.float)(unsafe.Pointer(&out[0])))
	}); err != nil {
		return nil, err
	}
	return out, nil
}

// Nearest — provider/online.go:63, called from serving.go:763.
func (t *b200Table) Nearest(feature, variant string, vector []float32, k int32) ([]string, error) {
	if len(vector) != t.dim || k < 0 {
		return nil, fferr.NewInvalidArgumentError(fmt.Errorf("expected vector of length %d and k >= 0", t.dim))
	}
	if k == 0 {
		return []string{}, nil
	}
	labels := make([]uint64, k)
	var count C.uint32_t
	// One goroutine per request, one vector per call (serving.go:744-771): the library's combining queue
	// coalesces the concurrent calls into batched launches, so no batching is needed on the Go side.
	runtime.LockOSThread()
	rc := C.ehb_index_search(t.ix, 1, (*C.float)(unsafe.Pointer(&vector[0])), C.uint32_t(k), 0,
		(*C.uint64_t)(unsafe.Pointer(&labels[0])), nil, &count)
	var msg string
	if rc != C.EHB_OK {
		msg = C.GoString(C.ehb_last_error())
	}
	runtime.UnlockOSThread()
	if rc != C.EHB_OK {
		return nil, fferr.NewResourceExecutionError(pt.B200VectorOnline.String(), feature, variant, fferr.ENTITY,
			fmt.Errorf("%s", msg))
	}
	t.mu.RLock()
	defer t.mu.RUnlock()
	out := make([]string, 0, int(count))
	for i := 0; i < int(count); i++ {
		out = append(out, t.keys[labels[i]])
	}
	return out, nil
}
