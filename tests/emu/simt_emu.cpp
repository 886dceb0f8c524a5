/*
 * tests/emu/simt_emu.cpp -- TEST INFRASTRUCTURE: the scheduler of the host SIMT emulator
 * declared in tests/emu/cuda_runtime.h (read that header first).
 */
#include "cuda_runtime.h"

#include <sys/mman.h>

namespace emu {

thread_local Block *blk = NULL;
int async_eager = 0;
int approx_ulp = 0;

[[noreturn]] void die(const char *what)
{
    Block *b = blk;
    if (b)
	fprintf(stderr, "simt_emu: %s (block %u of %u, thread %u of %u)\n", what, b->bidx.x, b->gdim.x,
		b->cur, b->bdim.x);
    else
	fprintf(stderr, "simt_emu: %s\n", what);
    abort();
}

static const size_t STACK_BYTES = 512u * 1024u;

static void trampoline()
{
    Block *b = blk;
    b->call(b->body);
    b = blk;
    Lane &l = b->lanes[b->cur];
    if (!l.queue.empty() || !l.groups.empty()) {
	/* copies still in flight when the thread exits would land in freed shared memory */
	for (size_t i = 0; i < l.groups.size(); i++)
	    if (l.groups[i])
		die("a thread exited with cp.async copies in flight");
	if (l.queue.size())
	    die("a thread exited with uncommitted cp.async copies");
	l.groups.clear();
    }
    l.done = true;
    b->progress++;
}

struct Worker {
    Block b;
    std::vector<void *> stacks;
    ~Worker()
    {
	for (size_t i = 0; i < stacks.size(); i++)
	    munmap(stacks[i], STACK_BYTES);
	free(b.smem);
    }
};

static void run_block(Worker &w, unsigned bx, unsigned grid, unsigned block, size_t smem, const void *body,
	void (*call)(const void *))
{
    Block &b = w.b;
    b.bidx = uint3{ bx, 0, 0 };
    b.bdim = uint3{ block, 1, 1 };
    b.gdim = uint3{ grid, 1, 1 };
    if (b.smem_bytes < smem + 64 || !b.smem) {
	free(b.smem);
	void *p = NULL;
	if (posix_memalign(&p, 1024, smem + 64) != 0)
	    die("out of memory");
	b.smem = (unsigned char *)p;
    }
    b.smem_bytes = smem;
    memset(b.smem, 0x5A, smem + 64);		/* shared memory starts out as garbage */
    b.lanes.resize(block);
    b.xchg.assign(block, 0);
    b.slots.assign((block + 31) / 32, std::vector<Slot>());
    for (size_t i = 0; i < b.slots.size(); i++)
	b.slots[i].reserve(64);		/* lanes wait on references into these */
    b.all = Slot{ 0, 0, 0 };
    b.progress = 0;
    b.body = body;
    b.call = call;
    while (w.stacks.size() < block) {
	void *s = mmap(NULL, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
	if (s == MAP_FAILED)
	    die("cannot map a fiber stack");
	w.stacks.push_back(s);
    }
    for (unsigned i = 0; i < block; i++) {
	Lane &l = b.lanes[i];
	l.done = false;
	l.queue.clear();
	l.groups.clear();
	getcontext(&l.ctx);
	l.ctx.uc_stack.ss_sp = w.stacks[i];
	l.ctx.uc_stack.ss_size = STACK_BYTES;
	l.ctx.uc_link = &b.sched;
	makecontext(&l.ctx, trampoline, 0);
    }
    blk = &b;
    unsigned remaining = block, idle = 0;
    while (remaining) {
	const unsigned long before = b.progress;
	for (unsigned i = 0; i < block; i++) {
	    if (b.lanes[i].done)
		continue;
	    b.cur = i;
	    swapcontext(&b.sched, &b.lanes[i].ctx);
	    if (b.lanes[i].done)
		remaining--;
	}
	if (b.progress == before) {
	    if (++idle > 2) {
		b.cur = 0;
		for (unsigned i = 0; i < block; i++)
		    if (!b.lanes[i].done) {
			b.cur = i;
			break;
		    }
		for (size_t wi = 0; wi < b.slots.size(); wi++)
		    for (size_t j = 0; j < b.slots[wi].size(); j++)
			fprintf(stderr, "simt_emu:   warp %zu mask %08x: %u of %d lanes waiting\n", wi,
				b.slots[wi][j].mask, b.slots[wi][j].count,
				__builtin_popcount(b.slots[wi][j].mask));
		fprintf(stderr, "simt_emu:   __syncthreads: %u of %u waiting; threads still running:", b.all.count,
			block);
		for (unsigned i = 0; i < block; i++)
		    if (!b.lanes[i].done)
			fprintf(stderr, " %u", i);
		fprintf(stderr, "\n");
		die("deadlock: every remaining thread waits in a *_sync that the others never reach");
	    }
	} else
	    idle = 0;
    }
    blk = NULL;
}

void run_grid(unsigned grid, unsigned block, size_t smem, const void *body, void (*call)(const void *))
{
    if (grid == 0 || block == 0)
	return;
    const char *e = getenv("FSK_EMU_ASYNC");
    async_eager = e && strcmp(e, "eager") == 0;
    approx_ulp = (e = getenv("FSK_EMU_ULP")) ? atoi(e) : 0;
    unsigned nthreads = std::thread::hardware_concurrency();
    if ((e = getenv("FSK_EMU_THREADS")))
	nthreads = (unsigned)atoi(e);
    nthreads = std::max(1u, std::min(std::min(nthreads, 8u), grid));
    std::atomic<unsigned> next(0);
    auto work = [&]() {
	Worker w;
	w.b.smem = NULL;
	w.b.smem_bytes = 0;
	for (;;) {
	    const unsigned bx = next.fetch_add(1);
	    if (bx >= grid)
		break;
	    run_block(w, bx, grid, block, smem, body, call);
	}
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; t++)
	pool.emplace_back(work);
    work();
    for (size_t t = 0; t < pool.size(); t++)
	pool[t].join();
}

}  /* namespace emu */
